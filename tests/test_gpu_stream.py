"""GPU tests of the ordered, back-pressured frame stream (nv12eq_stream_*) and of the
round-robin dispatcher on top of it.  Results are compared bit-exactly with the oracle through the C-ABI."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nv():
    import opencv_opencl_b200 as nv12eq
    nv12eq.build()
    return nv12eq


@pytest.fixture(scope="module")
def ctx(nv):
    c = nv.Context(device=0, max_width=3840, max_height=2160, slots=2)
    yield c
    c.close()


def test_stream_delivers_in_order_and_bit_exact(nv, ctx, oracle):
    W, H, n = 640, 360, 12
    frames = [oracle.c_synth_nv12(W, H, 2026, k) for k in range(n)]
    with nv.Stream(ctx, W, H, op=nv.OP_EQUALIZE, depth=4) as s:
        got, pushed = [], 0
        for k in range(n):
            if pushed - len(got) == 4:                 # keep the stream full without blocking ourselves
                got.append(s.pop())
            assert s.push(frames[k]) == k
            pushed += 1
        while len(got) < n:
            got.append(s.pop())
        assert s.pop(block=False) is None
        st = s.stats()
    assert [q for q, _ in got] == list(range(n))
    for q, f in got:
        assert np.array_equal(f, oracle.c_nv12_equalize_hist(frames[q], W, H)), q
    assert st["pushed"] == n and st["delivered"] == n and st["dropped_backpressure"] == 0 and st["max_in_flight"] == 4
    assert st["latency_us_max"] > 0


def test_stream_clahe_gray128_strided(nv, ctx, oracle):
    W, H, S = 320, 180, 384
    frames = [oracle.c_synth_nv12(W, H, 7, k, stride=S) for k in range(5)]
    with nv.Stream(ctx, W, H, op=nv.OP_CLAHE, stride=S, uv_mode=nv.UV_GRAY128, clip_limit=3.0, tiles=(4, 3), depth=2) as s:
        for k, f in enumerate(frames):
            assert s.push(f) == k
            q, out = s.pop()
            assert q == k
            want = oracle.c_nv12_clahe(f, W, H, 3.0, 4, 3, stride=S, uv_mode=oracle.UV_GRAY128, out=np.zeros_like(f))
            rows = out.reshape(-1, S)[:, :W]
            assert np.array_equal(rows, want.reshape(-1, S)[:, :W]), k


def test_stream_drop_newest_and_drop_oldest(nv, ctx, oracle):
    W, H = 256, 144
    frames = [oracle.c_synth_nv12(W, H, 11, k) for k in range(5)]
    # drop-newest: the 3rd..5th push find the queue full
    with nv.Stream(ctx, W, H, depth=2, full_policy=nv.FULL_DROP_NEWEST) as s:
        res = [s.push(f) for f in frames]
        assert res == [0, 1, None, None, None]
        a, b = s.pop(), s.pop()
        assert (a[0], b[0]) == (0, 1) and s.pop(block=False) is None
        assert np.array_equal(b[1], oracle.c_nv12_equalize_hist(frames[1], W, H))
        assert s.push(frames[0]) == 5                       # numbering continues past the dropped frames
        assert s.stats()["dropped_backpressure"] == 3
    # drop-oldest (GStreamer leaky=downstream, what the reference configures): the newest frames survive
    with nv.Stream(ctx, W, H, depth=2, full_policy=nv.FULL_DROP_OLDEST) as s:
        assert [s.push(f) for f in frames] == [0, 1, 2, 3, 4]
        a, b = s.pop(), s.pop()
        assert (a[0], b[0]) == (3, 4) and s.pop(block=False) is None
        assert np.array_equal(a[1], oracle.c_nv12_equalize_hist(frames[3], W, H))
        assert np.array_equal(b[1], oracle.c_nv12_equalize_hist(frames[4], W, H))
        assert s.stats()["dropped_backpressure"] == 3


def test_stream_blocking_producer_consumer_threads(nv, ctx, oracle):
    W, H, n = 640, 360, 48
    frames = [oracle.c_synth_nv12(W, H, 5, k) for k in range(8)]
    want = [oracle.c_nv12_clahe(f, W, H, 2.0, 8, 8) for f in frames]
    with nv.Stream(ctx, W, H, op=nv.OP_CLAHE, depth=4, full_policy=nv.FULL_BLOCK) as s:
        errors, got = [], []

        def producer():
            try:
                for k in range(n):
                    assert s.push(frames[k % 8]) == k
            except Exception as e:  # pragma: no cover
                errors.append(e)

        def consumer():
            try:
                for _ in range(n):
                    got.append(s.pop(block=True))
            except Exception as e:  # pragma: no cover
                errors.append(e)

        tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
        tp.start(); tc.start(); tp.join(60); tc.join(60)
        assert not errors and not tp.is_alive() and not tc.is_alive()
        st = s.stats()
    assert [q for q, _ in got] == list(range(n))
    assert all(np.array_equal(f, want[q % 8]) for q, f in got)
    assert st["delivered"] == n and st["max_in_flight"] <= 4


def test_stream_rejects_bad_arguments(nv, ctx):
    with pytest.raises(nv.Nv12eqError):
        nv.Stream(ctx, 64, 64, depth=0)
    with pytest.raises(nv.Nv12eqError):
        nv.Stream(ctx, 64, 64, op=nv.OP_CLAHE, tiles=(0, 8))
    with nv.Stream(ctx, 64, 64, depth=1) as s:
        short = np.zeros(10, np.uint8)
        with pytest.raises(nv.Nv12eqError) as e:
            s.push(short)
        assert e.value.status == nv.ERR_SHORT_BUFFER
        assert s.pop(block=False) is None


def test_frame_sharded_stream_over_two_contexts(nv, oracle):
    """Frame k -> context k mod 2 (two contexts on the one visible GPU stand in for two GPUs), capture order out."""
    W, H, n = 640, 360, 10
    frames = [oracle.c_synth_nv12(W, H, 3, k) for k in range(n)]
    ctxs = [nv.Context(0, W, H, 1) for _ in range(2)]
    fs = nv.sharding.FrameShardedStream([nv.Stream(c, W, H, depth=3) for c in ctxs])
    got = []
    for k in range(n):
        if fs.pending() == 6:
            got.append(fs.pop())
        assert fs.push(frames[k]) == k
    while fs.pending():
        got.append(fs.pop())
    fs.close()
    for c in ctxs:
        c.close()
    assert [k for k, _ in got] == list(range(n))
    assert all(np.array_equal(f, oracle.c_nv12_equalize_hist(frames[k], W, H)) for k, f in got)


def test_spatial_split_equalizer_single_rank(nv, ctx, oracle):
    """world size 1 degenerates to the whole frame; the 2-band exchange itself is covered on the CPU under gloo
    (tests/test_sharding.py) and stage by stage in test_gpu_parity.py::test_spatial_split_stage_api."""
    import torch
    W, H = 1920, 1080
    y = oracle.c_synth_nv12(W, H, 2026, 0)[:W * H]
    d_in = torch.from_numpy(y.copy()).cuda()
    d_out = torch.zeros_like(d_in)
    eq = nv.sharding.SpatialSplitEqualizer(ctx, W, H, rank=0, world=1)
    hist = eq.run(d_in, d_out, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(hist.cpu().numpy(), oracle.c_hist256(y.reshape(H, W)))
    assert np.array_equal(d_out.cpu().numpy().reshape(H, W), oracle.c_equalize_hist(y.reshape(H, W)))


def _spatial_split_worker(rank, world, port, q):
    """one rank of the 2-GPU spatial split: own band of one 4K luma plane, NCCL all-reduce of the 256-bin histogram"""
    import torch
    import torch.distributed as dist
    import opencv_opencl_b200 as nv
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        W, H = 3840, 2160
        y = O.c_synth_nv12(W, H, 2026, 3)[:W * H].reshape(H, W)
        want = O.c_equalize_hist(y)
        ok = True
        with nv.Context(rank, W, H, 1) as ctx:
            eq = nv.sharding.SpatialSplitEqualizer(ctx, W, H, rank, world)
            first, rows = eq.band
            d_in = torch.from_numpy(np.ascontiguousarray(y[first:first + rows]).reshape(-1)).cuda()
            for stream in (None, torch.cuda.Stream()):       # torch's current stream, and a side stream
                d_out = torch.zeros_like(d_in)
                for _ in range(3):                           # repeated: a racing zero-fill would corrupt a later run
                    hist = eq.run(d_in, d_out, stream=stream)
                torch.cuda.synchronize()
                ok = ok and np.array_equal(hist.cpu().numpy(), O.c_hist256(y))
                ok = ok and np.array_equal(d_out.cpu().numpy().reshape(rows, W), want[first:first + rows])
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))


def test_spatial_split_equalizer_two_gpus(nv):
    """SURVEY 8e optional mode with a real exchange: two ranks, two GPUs, one NCCL all-reduce per frame."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctxmp.Process(target=_spatial_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res


def test_spatial_split_clahe_stage_api_and_single_rank(nv, ctx, oracle):
    """The band stages of the spatially split CLAHE on ONE GPU: two bands processed one after the other with the halo rows copied
    by hand (what the NCCL send/recv does between ranks), and the world-size-1 form of SpatialSplitClahe."""
    import torch
    for (W, H, tiles, clip, splits) in ((1920, 1080, (8, 8), 2.0, [(0, 4), (4, 4)]), (3840, 2160, (8, 8), 3.0, [(0, 3), (3, 3), (6, 2)]),
                                        (640, 480, (4, 6), 40.0, [(0, 1), (1, 5)])):
        tx, ty = tiles
        th = H // ty
        y = oracle.c_synth_nv12(W, H, 2026, 1)[:W * H].reshape(H, W)
        want = oracle.c_clahe(y, clip, tx, ty)
        d_y = torch.from_numpy(y.copy()).cuda()
        st = torch.cuda.current_stream()
        all_luts = torch.zeros(ty * tx * 256, dtype=torch.uint8, device="cuda")
        for first, nb in splits:   # stage 1 on every band
            ctx.clahe_band_luts_device(d_y[first * th:], W, H, clip, tiles, first, nb, all_luts[first * tx * 256:], stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(all_luts.cpu().numpy().reshape(-1, 256), oracle.c_clahe_tile_luts(y, clip, tx, ty))
        d_out = torch.zeros_like(d_y)
        for first, nb in splits:   # the exchange by hand, then stage 2
            halo = torch.full(((nb + 2) * tx * 256,), 0xEE, dtype=torch.uint8, device="cuda")   # rows outside the frame stay garbage
            lo, hi = max(first - 1, 0), min(first + nb + 1, ty)
            halo[(lo - (first - 1)) * tx * 256:(hi - (first - 1)) * tx * 256] = all_luts[lo * tx * 256:hi * tx * 256]
            ctx.clahe_band_apply_device(d_y[first * th:], d_out[first * th:], W, H, tiles, first, nb, halo, stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want), (W, H)
        one = nv.sharding.SpatialSplitClahe(ctx, W, H, clip, tiles, rank=0, world=1)
        d_out.zero_()
        one.run(d_y.reshape(-1), d_out.reshape(-1), stream=st)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)
    with pytest.raises(nv.Nv12eqError):   # a grid that does not divide the frame is refused
        ctx.clahe_band_luts_device(d_y, 1918, 1078, 2.0, (8, 8), 0, 4, all_luts)


def _spatial_clahe_worker(rank, world, port, q):
    """one rank of the 2-GPU spatially split CLAHE: own tile rows of one 4K luma plane, NCCL send/recv of one LUT row"""
    import torch
    import torch.distributed as dist
    import opencv_opencl_b200 as nv
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        W, H, tiles, clip = 3840, 2160, (8, 8), 2.0
        y = O.c_synth_nv12(W, H, 2026, 3)[:W * H].reshape(H, W)
        want = O.c_clahe(y, clip, *tiles)
        ok = True
        with nv.Context(rank, W, H, 1) as ctx:
            sp = nv.sharding.SpatialSplitClahe(ctx, W, H, clip, tiles, rank, world)
            first, rows = sp.band
            d_in = torch.from_numpy(np.ascontiguousarray(y[first:first + rows]).reshape(-1)).cuda()
            for stream in (None, torch.cuda.Stream()):
                d_out = torch.zeros_like(d_in)
                for _ in range(3):
                    sp.run(d_in, d_out, stream=stream)
                torch.cuda.synchronize()
                ok = ok and np.array_equal(d_out.cpu().numpy().reshape(rows, W), want[first:first + rows])
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))


def test_spatial_split_clahe_two_gpus(nv):
    """SURVEY 8e optional mode for CLAHE with a real exchange: two ranks, two GPUs, one LUT-row send/recv per neighbour."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctxmp.Process(target=_spatial_clahe_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res


def test_cpp_example_runs_both_modes(nv):
    """The C++ example (reference worker loop over the C-ABI): worker pool + reorder buffer, and the nv12eq_stream form."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ex = os.path.join(root, "examples")
    subprocess.run(["make", "-C", ex, "-B", "-s"], check=True)
    for extra in (["--workers", "3"], ["--stream"], ["--op", "clahe", "--stream"], ["--op", "clahe", "--workers", "2"]):
        out = subprocess.run([os.path.join(ex, "worker_demo"), "--frames", "24", "--width", "640", "--height", "360"] + extra,
                             capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "processed 24, delivered in order 24, out of order 0, errors 0" in out.stdout, out.stdout


def test_sharded_batch_over_contexts(nv, oracle):
    """One process, N contexts (one per visible GPU; two on the same GPU when only one is visible): contiguous slices of a
    host batch, no collective, results identical to the single-context batch."""
    import torch
    W, H, n = 640, 360, 11
    devs = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    ctxs = [nv.Context(d, W, H, 2) for d in devs]
    sb = nv.sharding.ShardedBatch(ctxs)
    frames = np.stack([oracle.c_synth_nv12(W, H, 9, k) for k in range(n)])
    pitch = frames.shape[1]
    out = np.zeros_like(frames)
    sb.equalize_hist(frames, W, H, out, n, pitch)
    assert np.array_equal(out, oracle.c_nv12_batch("equalize", frames, W, H))
    out2 = np.zeros_like(frames)
    sb.clahe(frames, W, H, out2, n, pitch, clip_limit=2.0, tiles=(8, 8), uv_mode=nv.UV_GRAY128)
    assert np.array_equal(out2, oracle.c_nv12_batch("clahe", frames, W, H, clip=2.0, tx=8, ty=8, uv_mode=oracle.UV_GRAY128))
    sb.close()
    for c in ctxs:
        c.close()
