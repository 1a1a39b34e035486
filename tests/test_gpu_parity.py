"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C-ABI of libnv12eq.so,
against the CPU oracle and the committed cv2-generated golden vectors.  Bit-exact everywhere (byte/integer work;
the fp32 steps reproduce OpenCV's separately rounded operations, so the tolerance is 0)."""
import hashlib

import numpy as np
import pytest

from cases import dist_image

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def nv():
    import opencv_opencl_b200 as nv12eq
    nv12eq.build()
    return nv12eq


@pytest.fixture(scope="module")
def ctx(nv):
    c = nv.Context(device=0, max_width=8192, max_height=4608, slots=2)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


# ------------------------------------------------------------------------------------------------------------
# host frame-in / frame-out against golden digests (cv2) and the oracle
# ------------------------------------------------------------------------------------------------------------
def test_equalize_golden_synth(ctx, oracle, golden):
    for rec in golden["synth"]:
        W, H = rec["W"], rec["H"]
        nv12 = oracle.c_synth_nv12(W, H, rec["seed"], rec["frame"])
        out = ctx.equalize_hist(nv12, W, H)
        assert sha(out[:W * H]) == rec["eq"], (W, H, rec["frame"])
        assert np.array_equal(out[W * H:], nv12[W * H:]), "UV passthrough"


def test_clahe_golden_synth(ctx, oracle, golden):
    n = 0
    for rec in golden["synth"]:
        W, H = rec["W"], rec["H"]
        nv12 = oracle.c_synth_nv12(W, H, rec["seed"], rec["frame"])
        for key, digest in rec["clahe"].items():
            clip, tx, ty = key.split(":")
            out = ctx.clahe(nv12, W, H, float(clip), (int(tx), int(ty)))
            assert sha(out[:W * H]) == digest, (W, H, key)
            assert np.array_equal(out[W * H:], nv12[W * H:])
            n += 1
    assert n > 100


def test_small_distributions_golden(ctx, oracle, golden):
    """Tiny and degenerate planes (1x1, constant, binary, one-hot): NV12 needs a chroma plane, so build one."""
    for rec in golden["dist"]:
        W, H = rec["W"], rec["H"]
        y = dist_image(rec["kind"], W, H, rec["seed"])
        nv12 = np.concatenate([y.reshape(-1), np.full(W * (H // 2), 77, np.uint8)])
        out = ctx.equalize_hist(nv12, W, H)
        assert sha(out[:W * H].reshape(H, W)) == rec["eq"], rec
        assert (out[W * H:] == 77).all()
        for key, digest in rec["clahe"].items():
            clip, tx, ty = key.split(":")
            out = ctx.clahe(nv12, W, H, float(clip), (int(tx), int(ty)))
            assert sha(out[:W * H].reshape(H, W)) == digest, (rec, key)


def test_raw_fixtures(ctx, fixtures):
    for base in [k[:-3] for k in fixtures if k.endswith("_in") and not k.startswith("color")]:
        y = fixtures[base + "_in"]
        H, W = y.shape
        nv12 = np.concatenate([y.reshape(-1), np.zeros(W * (H // 2), np.uint8)])
        assert np.array_equal(ctx.equalize_hist(nv12, W, H)[:W * H].reshape(H, W), fixtures[base + "_eq"]), base
        for k in fixtures:
            if k.startswith(base + "_clahe_"):
                clip, tx, ty = k[len(base) + 7:].split("_")
                got = ctx.clahe(nv12, W, H, float(clip), (int(tx), int(ty)))[:W * H].reshape(H, W)
                assert np.array_equal(got, fixtures[k]), k


def test_reference_shaped_operators(nv, ctx, oracle):
    """equalizeHist / createCLAHE(...).apply keep the reference's names and argument meaning."""
    W, H = 1280, 720
    nv12 = oracle.c_synth_nv12(W, H, 2026, 0)
    assert np.array_equal(nv.equalizeHist(nv12, W, H, ctx=ctx), oracle.c_nv12_equalize_hist(nv12, W, H))
    clahe = nv.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8), ctx=ctx)
    assert np.array_equal(clahe.apply(nv12, W, H, uv_mode=nv.UV_GRAY128),
                          oracle.c_nv12_clahe(nv12, W, H, 2.0, 8, 8, uv_mode=oracle.UV_GRAY128))
    clahe.setClipLimit(3.0)
    clahe.setTilesGridSize((4, 4))
    assert np.array_equal(clahe.apply(nv12, W, H), oracle.c_nv12_clahe(nv12, W, H, 3.0, 4, 4))


@pytest.mark.parametrize("W,H,S", [(100, 36, 128), (1918, 1078, 1920), (70, 38, 71), (640, 360, 1024), (33, 17, 48)])
def test_strided_frames_and_uv_modes(nv, ctx, oracle, W, H, S):
    nv12 = oracle.c_synth_nv12(W, H, 2026, 3, stride=S)
    for uv_mode in (nv.UV_COPY, nv.UV_GRAY128, nv.UV_SKIP):
        pre = np.full_like(nv12, 9)
        want = oracle.c_nv12_equalize_hist(nv12, W, H, stride=S, uv_mode=uv_mode, out=pre.copy())
        got = ctx.equalize_hist(nv12, W, H, stride=S, uv_mode=uv_mode, out=pre.copy())
        assert np.array_equal(got, want), ("eq", uv_mode)  # includes: padding bytes and skipped chroma untouched
        want = oracle.c_nv12_clahe(nv12, W, H, 2.0, 4, 3, stride=S, uv_mode=uv_mode, out=pre.copy())
        got = ctx.clahe(nv12, W, H, 2.0, (4, 3), stride=S, uv_mode=uv_mode, out=pre.copy())
        assert np.array_equal(got, want), ("clahe", uv_mode)


def test_in_place_and_pinned(nv, ctx, oracle):
    W, H = 1920, 1080
    nv12 = oracle.c_synth_nv12(W, H, 2026, 1)
    want_eq = oracle.c_nv12_equalize_hist(nv12, W, H)
    want_cl = oracle.c_nv12_clahe(nv12, W, H)
    buf = nv12.copy()
    ctx.equalize_hist(buf, W, H, out=buf)
    assert np.array_equal(buf, want_eq)
    pin_in, pin_out = nv.PinnedBuffer(nv12.nbytes), nv.PinnedBuffer(nv12.nbytes)
    pin_in.array[:] = nv12
    ctx.clahe(pin_in.array, W, H, out=pin_out.array)
    assert np.array_equal(pin_out.array, want_cl)
    ctx.clahe(pin_in.array, W, H, out=pin_in.array)
    assert np.array_equal(pin_in.array, want_cl)
    pin_in.free(); pin_out.free()


def test_host_batches_and_async_slots(nv, ctx, oracle):
    W, H, n = 1280, 720, 13
    frames = np.stack([oracle.c_synth_nv12(W, H, 2026, k) for k in range(n)])
    want_eq = oracle.c_nv12_batch("equalize", frames, W, H)
    want_cl = oracle.c_nv12_batch("clahe", frames, W, H, clip=2.0, tx=8, ty=8, uv_mode=oracle.UV_GRAY128)
    assert np.array_equal(ctx.equalize_hist_batch(frames, W, H), want_eq)
    assert np.array_equal(ctx.clahe_batch(frames, W, H, 2.0, (8, 8), uv_mode=nv.UV_GRAY128), want_cl)
    # two slots in flight, different ops
    out0, out1 = np.empty_like(frames[:6]), np.empty_like(frames[6:])
    pitch = frames.shape[1]
    ctx.submit_equalize_hist(0, frames[:6], out0, 6, pitch, W, H)
    ctx.submit_clahe(1, frames[6:], out1, n - 6, pitch, W, H, 2.0, (8, 8), uv_mode=nv.UV_GRAY128)
    with pytest.raises(nv.Nv12eqError):  # slot 0 is busy
        ctx.submit_equalize_hist(0, frames[:6], out0, 6, pitch, W, H)
    ctx.wait(1); ctx.wait(0)
    assert ctx.query(0) and ctx.query(1)
    assert np.array_equal(out0, want_eq[:6]) and np.array_equal(out1, want_cl[6:])


def test_error_behaviour(nv, ctx, oracle):
    """Status codes instead of exceptions/aborts; the reference drops such frames (OpenCVequalHist.cpp:132-137)."""
    W, H = 64, 48
    nv12 = oracle.c_synth_nv12(W, H, 2026, 0)
    out = np.empty_like(nv12)
    assert ctx.equalize_hist(nv12[:-1], W, H, out=out, raw_status=True) == nv.ERR_SHORT_BUFFER
    assert ctx.equalize_hist(nv12, W, H, out=out[:-1], raw_status=True) == nv.ERR_SHORT_BUFFER
    assert ctx.equalize_hist(nv12, 0, H, out=out, raw_status=True) == nv.ERR_INVALID_ARGUMENT
    assert ctx.equalize_hist(nv12, W, H, stride=W - 1, out=out, raw_status=True) == nv.ERR_INVALID_ARGUMENT
    assert ctx.equalize_hist(nv12, W, H, uv_mode=5, out=out, raw_status=True) == nv.ERR_INVALID_ARGUMENT
    assert ctx.clahe(nv12, W, H, 2.0, (0, 8), out=out, raw_status=True) == nv.ERR_INVALID_ARGUMENT
    assert ctx.equalize_hist(nv12, 10000, 10, out=out, raw_status=True) == nv.ERR_TOO_LARGE
    assert "exceeds" in ctx.last_error()
    overlapping = np.zeros(nv12.nbytes + 16, np.uint8)
    assert ctx.equalize_hist(overlapping[:nv12.nbytes], W, H, out=overlapping[16:], raw_status=True) == nv.ERR_INVALID_ARGUMENT
    errs = ctx.counters()["errors"]
    assert errs >= 8
    # the context still works afterwards
    assert np.array_equal(ctx.equalize_hist(nv12, W, H), oracle.c_nv12_equalize_hist(nv12, W, H))
    with pytest.raises(nv.Nv12eqError):
        nv.Context(device=99)


# ------------------------------------------------------------------------------------------------------------
# device-resident forms
# ------------------------------------------------------------------------------------------------------------
def test_synth_device_matches_appendix_b(ctx, oracle, torch):
    for (W, H, S, n, f0) in [(1920, 1080, 1920, 3, 0), (3840, 2160, 3840, 2, 0), (1918, 1078, 1918, 2, 5), (100, 36, 128, 2, 1)]:
        pitch = S * (H + H // 2)
        d = torch.zeros(n * pitch, dtype=torch.uint8, device="cuda")
        ctx.synth_nv12_device(d, n, pitch, W, H, stride=S, seed=2026, first_frame=f0, stream=torch.cuda.current_stream())
        got = d.cpu().numpy().reshape(n, pitch)
        for k in range(n):
            assert np.array_equal(got[k], oracle.c_synth_nv12(W, H, 2026, f0 + k, stride=S)), (W, H, k)
    W, H = 322, 200
    d = torch.zeros(2 * W * H * 3, dtype=torch.uint8, device="cuda")
    ctx.synth_bgr_device(d, 2, W * H * 3, W, H, first_frame=0, stream=torch.cuda.current_stream())
    got = d.cpu().numpy().reshape(2, H, W, 3)
    assert np.array_equal(got[0], oracle.c_synth_bgr(W, H, 0)) and np.array_equal(got[1], oracle.c_synth_bgr(W, H, 1))


@pytest.mark.parametrize("tuning", [dict(), dict(schedule=2), dict(lag_frames=-1), dict(lag_frames=1, chunks_per_frame=7),
                                    dict(lag_frames=5, chunks_per_frame=301, ctas_per_sm=1), dict(ctas_per_sm=2, chunks_per_frame=1)])
def test_device_batch_all_schedules(nv, oracle, torch, tuning):
    """Every schedule / chunking / lag gives the same bytes (the ticket-lag scheme is only a schedule)."""
    W, H, n = 1280, 720, 9
    pitch = nv.nv12_frame_bytes(W, H)
    frames = np.stack([oracle.c_synth_nv12(W, H, 2026, k) for k in range(n)])
    want_eq = oracle.c_nv12_batch("equalize", frames, W, H)
    want_cl = oracle.c_nv12_batch("clahe", frames, W, H, clip=2.0, tx=8, ty=8)
    with nv.Context(0, W, H, 1) as c:
        c.set_tuning(**tuning)
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.empty_like(d_in)
        st = torch.cuda.current_stream()
        for rep in range(3):  # repeated launches reuse the self-cleaning workspace
            d_out.zero_()
            c.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
            assert np.array_equal(d_out.cpu().numpy(), want_eq), (tuning, rep)
            d_out.zero_()
            c.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (8, 8), stream=st)
            assert np.array_equal(d_out.cpu().numpy(), want_cl), (tuning, rep)
        # in place on the device
        d_io = d_in.clone()
        c.equalize_hist_device(d_io, d_io, n, pitch, W, H, stream=st)
        assert np.array_equal(d_io.cpu().numpy(), want_eq)
        d_io = d_in.clone()
        c.clahe_device(d_io, d_io, n, pitch, W, H, 2.0, (8, 8), stream=st)
        assert np.array_equal(d_io.cpu().numpy(), want_cl)


def _lut_properties(torch, d_in, d_out, n, W, H):
    """Size-independent properties of equalizeHist on the device: the map is a per-frame LUT (same input value ->
    same output value), the LUT is monotone non-decreasing, the smallest occupied bin maps to 0 and the largest to 255,
    and chroma is untouched."""
    pitch = W * (H + H // 2)
    for k in range(n):
        yi = d_in[k * pitch:k * pitch + W * H].long()
        yo = d_out[k * pitch:k * pitch + W * H].long()
        lo = torch.full((256,), 256, device="cuda", dtype=torch.long).scatter_reduce(0, yi, yo, "amin")
        hi = torch.full((256,), -1, device="cuda", dtype=torch.long).scatter_reduce(0, yi, yo, "amax")
        occ = hi >= 0
        assert bool((lo[occ] == hi[occ]).all()), "not a LUT"
        vals = hi[occ]
        assert bool((vals[1:] >= vals[:-1]).all()), "LUT not monotone"
        assert int(vals[0]) == 0 and int(vals[-1]) == 255
        assert torch.equal(d_in[k * pitch + W * H:(k + 1) * pitch], d_out[k * pitch + W * H:(k + 1) * pitch])


def test_full_size_config2_equalize_4k_batch(nv, ctx, oracle, golden, torch):
    """BASELINE config 2: equalizeHist on a 256-frame 4K NV12 batch, device resident."""
    W, H, n = 3840, 2160, 256
    pitch = nv.nv12_frame_bytes(W, H)
    st = torch.cuda.current_stream()
    d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros_like(d_in)
    ctx.synth_nv12_device(d_in, n, pitch, W, H, seed=2026, first_frame=0, stream=st)
    ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
    torch.cuda.synchronize()
    by_frame = {r["frame"]: r for r in golden["synth"] if (r["W"], r["H"]) == (W, H)}
    for k in (0, 1):  # golden digests (cv2)
        assert sha(d_out[k * pitch:k * pitch + W * H].cpu().numpy()) == by_frame[k]["eq"]
    for k in (2, 100, 255):  # oracle on the same seeded inputs
        frame = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
        assert np.array_equal(frame, oracle.c_synth_nv12(W, H, 2026, k))
        assert np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), oracle.c_nv12_equalize_hist(frame, W, H))
    _lut_properties(torch, d_in, d_out, n, W, H)
    # checksum of checksums against a second run with a different schedule
    ref_sum = d_out.view(n, pitch).long().sum(dim=1)
    with nv.Context(0, W, H, 1) as c2:
        c2.set_tuning(schedule=2)
        d_out2 = torch.zeros_like(d_in)
        c2.equalize_hist_device(d_in, d_out2, n, pitch, W, H, stream=st)
        assert torch.equal(d_out2.view(n, pitch).long().sum(dim=1), ref_sum)
        assert torch.equal(d_out2, d_out)


def test_full_size_clahe_1080p_and_4k(nv, ctx, oracle, golden, torch):
    """BASELINE configs 3/4: CLAHE clip 2.0, 8x8 on 1080p and 4K streams (device resident batches)."""
    st = torch.cuda.current_stream()
    for (W, H, n, probes) in [(1920, 1080, 256, (2, 127, 255)), (3840, 2160, 256, (2, 31, 128, 255))]:  # the bench's batch sizes
        pitch = nv.nv12_frame_bytes(W, H)
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros_like(d_in)
        ctx.synth_nv12_device(d_in, n, pitch, W, H, seed=2026, first_frame=0, stream=st)
        ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (8, 8), stream=st)
        torch.cuda.synchronize()
        by_frame = {r["frame"]: r for r in golden["synth"] if (r["W"], r["H"]) == (W, H)}
        for k in (0, 1):
            assert sha(d_out[k * pitch:k * pitch + W * H].cpu().numpy()) == by_frame[k]["clahe"]["2.0:8:8"]
        for k in probes:
            frame = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
            assert np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), oracle.c_nv12_clahe(frame, W, H, 2.0, 8, 8))
        v = d_in.view(n, pitch)[:, W * H:]
        assert torch.equal(v, d_out.view(n, pitch)[:, W * H:]), "chroma passthrough"
        # determinism / schedule independence
        d_out2 = torch.zeros_like(d_in)
        ctx.clahe_device(d_in, d_out2, n, pitch, W, H, 2.0, (8, 8), stream=st)
        assert torch.equal(d_out, d_out2)
        del d_in, d_out, d_out2


def test_spatial_split_stage_api(nv, ctx, oracle, torch):
    """SURVEY 8e optional mode on one GPU: two row bands -> two partial histograms -> sum (what ncclAllReduce would do)
    -> every band applies the LUT of the summed histogram.  Result == whole-frame equalizeHist."""
    W, H = 1920, 1080
    nv12 = oracle.c_synth_nv12(W, H, 2026, 4)
    want = oracle.c_equalize_hist(nv12[:W * H].reshape(H, W))
    d_y = torch.from_numpy(nv12[:W * H].copy()).cuda()
    d_o = torch.zeros_like(d_y)
    st = torch.cuda.current_stream()
    bands = [(0, 500), (500, H)]
    hists = []
    for (r0, r1) in bands:
        h = torch.zeros(256, dtype=torch.int32, device="cuda")
        ctx.hist_device(d_y[r0 * W:], 1, 0, W, r1 - r0, h, stream=st)
        hists.append(h)
    assert np.array_equal((hists[0] + hists[1]).cpu().numpy(), np.bincount(nv12[:W * H], minlength=256))
    total = (hists[0] + hists[1]).contiguous()
    for (r0, r1) in bands:
        ctx.equalize_apply_device(d_y[r0 * W:], d_o[r0 * W:], 1, 0, W, r1 - r0, total, W * H, stream=st)
    assert np.array_equal(d_o.cpu().numpy().reshape(H, W), want)


# ------------------------------------------------------------------------------------------------------------
# colour path
# ------------------------------------------------------------------------------------------------------------
def test_color_path_golden(nv, ctx, oracle, golden, fixtures):
    for rec in golden["color"]:
        W, H = rec["W"], rec["H"]
        bgr = oracle.c_synth_bgr(W, H, 0)
        for name, mode in (("yuv", nv.COLOR_YUV), ("ycrcb", nv.COLOR_YCRCB)):
            assert sha(ctx.color_equalize(bgr, mode)) == rec[f"{name}_eq"], (W, H, name)
            assert sha(ctx.color_clahe(bgr, 3.0, (4, 4), mode)) == rec[f"{name}_clahe_3.0_4_4"], (W, H, name)
    bgr = fixtures["color_31x9_in"]
    assert np.array_equal(ctx.color_equalize(bgr, nv.COLOR_YUV), fixtures["color_31x9_yuv_eq"])
    assert np.array_equal(ctx.color_equalize(bgr, nv.COLOR_YCRCB), fixtures["color_31x9_ycrcb_eq"])


def test_color_device_batch_and_stride(nv, ctx, oracle, torch):
    W, H, n = 322, 200, 3
    frames = np.stack([oracle.c_synth_bgr(W, H, k) for k in range(n)])
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros_like(d_in)
    ctx.color_equalize_device(d_in, d_out, n, W * H * 3, W, H, stream=torch.cuda.current_stream())
    got = d_out.cpu().numpy()
    for k in range(n):
        assert np.array_equal(got[k], oracle.c_color_equalize(frames[k], oracle.COLOR_YUV))
    # strided host image (a view into a wider buffer)
    wide = np.zeros((H, W + 10, 3), np.uint8)
    wide[:, :W] = frames[0]
    out = np.full_like(wide, 5)
    ctx.color_equalize(wide[:, :W], nv.COLOR_YCRCB, out=out[:, :W])
    assert np.array_equal(out[:, :W], oracle.c_color_equalize(frames[0], oracle.COLOR_YCRCB))
    assert (out[:, W:] == 5).all()


def test_color_full_bgr_cube(nv, oracle):
    """All 2^24 BGR values (a 4096x4096 image) through the fused two-pass colour equalization, both conversions: every
    saturation corner of the Q14 forward / inverse formulas is hit.  Also a size whose pixel count is not a multiple of
    512 (ragged last warp round) and an in-place call."""
    cube = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([(cube & 255), (cube >> 8) & 255, (cube >> 16) & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    with nv.Context(0, 4096, 4096, 1) as c:
        for mode, omode in ((nv.COLOR_YUV, oracle.COLOR_YUV), (nv.COLOR_YCRCB, oracle.COLOR_YCRCB)):
            assert np.array_equal(c.color_equalize(bgr, mode), oracle.c_color_equalize(bgr, omode)), mode
        small = oracle.c_synth_bgr(328, 202, 3)            # 66256 pixels = 129 rounds + 208 pixels
        want = oracle.c_color_equalize(small, oracle.COLOR_YUV)
        assert np.array_equal(c.color_equalize(small, nv.COLOR_YUV), want)
        buf = small.copy()
        c.color_equalize(buf, nv.COLOR_YUV, out=buf)
        assert np.array_equal(buf, want)


def test_randomized_geometries(nv, oracle):
    """Seeded sweep over ragged geometries: odd widths / heights, strides with every alignment, tile grids that do not
    divide the image (OpenCV's reflect-101 padding path), grids larger than the image, tiny frames, all uv modes, in-place
    and batched calls.  Every result is compared bit-exactly with the oracle."""
    rng = np.random.default_rng(20261018)
    with nv.Context(0, 1024, 1024, 2) as c:
        for case in range(70):
            W = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 63, 64, 65, 120, 127, 128, 130, 240, 250, 256, 333, 480, 511, 640]))
            H = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 17, 30, 33, 64, 67, 100, 135, 136, 270, 271]))
            S = W + int(rng.choice([0, 0, 1, 3, 5, 8, 16, 29, 32]))
            uv_mode = int(rng.integers(0, 3))
            nv12 = oracle.c_synth_nv12(W, H, int(rng.integers(1, 1 << 20)), int(rng.integers(0, 100)), stride=S)
            if rng.random() < 0.2:                       # flat / two-valued frames: constant-image and i0 edge cases
                nv12[:] = int(rng.integers(0, 256))
                if rng.random() < 0.5 and nv12.size > 3:
                    nv12[int(rng.integers(0, S * H))] = int(rng.integers(0, 256))
            pre = rng.integers(0, 256, nv12.size, dtype=np.uint8)
            got = c.equalize_hist(nv12, W, H, stride=S, uv_mode=uv_mode, out=pre.copy())
            want = oracle.c_nv12_equalize_hist(nv12, W, H, stride=S, uv_mode=uv_mode, out=pre.copy())
            assert np.array_equal(got, want), ("eq", case, W, H, S, uv_mode)
            tx, ty = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16])), int(rng.choice([1, 2, 3, 4, 6, 8, 9, 16]))
            clip = float(rng.choice([0.0, 0.5, 1.0, 2.0, 3.0, 40.0]))
            got = c.clahe(nv12, W, H, clip, (tx, ty), stride=S, uv_mode=uv_mode, out=pre.copy())
            want = oracle.c_nv12_clahe(nv12, W, H, clip, tx, ty, stride=S, uv_mode=uv_mode, out=pre.copy())
            assert np.array_equal(got, want), ("clahe", case, W, H, S, uv_mode, clip, tx, ty)
            if case % 5 == 0:                            # in place, and a 3-frame batch of the same geometry
                buf = nv12.copy()
                c.clahe(buf, W, H, clip, (tx, ty), stride=S, uv_mode=nv.UV_COPY, out=buf)
                assert np.array_equal(buf, oracle.c_nv12_clahe(nv12, W, H, clip, tx, ty, stride=S, out=nv12.copy())), ("inplace", case)
                batch = np.stack([nv12, np.roll(nv12, 7), nv12[::-1].copy()])
                outb = c.equalize_hist_batch(batch, W, H, stride=S, out=np.zeros_like(batch))
                for k in range(3):
                    wantk = oracle.c_nv12_equalize_hist(batch[k], W, H, stride=S, out=np.zeros_like(nv12))
                    assert np.array_equal(outb[k].reshape(-1, S)[:, :W], wantk.reshape(-1, S)[:, :W]), ("batch", case, k)


def test_large_frames_many_frames_and_fine_grids(nv, oracle, torch):
    """Size extremes: an 8K luma plane (33 M pixels, 518 K-pixel tiles, 960-pixel cells), a 64x64 tile grid on 1080p (tiny
    30x17 tiles, 4225 interpolation cells), a 1x1 grid, and a 1500-frame batch of small frames in one launch."""
    with nv.Context(0, 7680, 4320, 2) as c:
        W, H = 7680, 4320
        f8k = oracle.c_synth_nv12(W, H, 2026, 3)
        assert np.array_equal(c.equalize_hist(f8k, W, H), oracle.c_nv12_equalize_hist(f8k, W, H))
        assert np.array_equal(c.clahe(f8k, W, H, 2.0, (8, 8)), oracle.c_nv12_clahe(f8k, W, H, 2.0, 8, 8))
        W, H = 1920, 1080
        f = oracle.c_synth_nv12(W, H, 2026, 5)
        for (tx, ty) in ((64, 64), (1, 1), (37, 3)):
            assert np.array_equal(c.clahe(f, W, H, 4.0, (tx, ty)), oracle.c_nv12_clahe(f, W, H, 4.0, tx, ty)), (tx, ty)
        W, H, n = 64, 48, 1500
        pitch = nv.nv12_frame_bytes(W, H)
        st = torch.cuda.current_stream()
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
        d_eq, d_cl = torch.zeros_like(d_in), torch.zeros_like(d_in)
        c.synth_nv12_device(d_in, n, pitch, W, H, seed=2026, first_frame=0, stream=st)
        c.equalize_hist_device(d_in, d_eq, n, pitch, W, H, stream=st)
        c.clahe_device(d_in, d_cl, n, pitch, W, H, 2.0, (8, 8), stream=st)
        torch.cuda.synchronize()
        h_in, h_eq, h_cl = d_in.cpu().numpy().reshape(n, pitch), d_eq.cpu().numpy().reshape(n, pitch), d_cl.cpu().numpy().reshape(n, pitch)
        assert np.array_equal(h_eq, oracle.c_nv12_batch("equalize", h_in, W, H))
        assert np.array_equal(h_cl, oracle.c_nv12_batch("clahe", h_in, W, H, clip=2.0, tx=8, ty=8))


def test_color_host_batch(nv, ctx, oracle):
    """nv12eq_color_equalize_batch: several BGR frames per call, pipelined over the slots; pinned and pageable memory."""
    W, H, n = 640, 360, 7
    frames = np.stack([oracle.c_synth_bgr(W, H, k) for k in range(n)])
    want = np.stack([oracle.c_color_equalize(frames[k], oracle.COLOR_YCRCB) for k in range(n)])
    assert np.array_equal(ctx.color_equalize_batch(frames, nv.COLOR_YCRCB), want)
    pin_in, pin_out = nv.PinnedBuffer(frames.nbytes), nv.PinnedBuffer(frames.nbytes)
    pin_in.array[:] = frames.reshape(-1)
    ctx.color_equalize_batch(pin_in.array.reshape(frames.shape), nv.COLOR_YCRCB, out=pin_out.array.reshape(frames.shape))
    assert np.array_equal(pin_out.array.reshape(frames.shape), want)
    pin_in.free(); pin_out.free()


def test_device_calls_alternating_streams(nv, ctx, oracle):
    """The device forms of one context share a workspace; consecutive calls on different CUDA streams must be ordered by the
    library (the later stream waits for the earlier one), not race on the ticket counter and the tables."""
    import torch
    W, H, n = 640, 360, 24
    pitch = nv.nv12_frame_bytes(W, H)
    frames = np.stack([oracle.c_synth_nv12(W, H, 77, k) for k in range(n)])
    d_in = torch.from_numpy(frames).cuda()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [(torch.zeros_like(d_in), torch.zeros_like(d_in)) for _ in range(6)]
    torch.cuda.synchronize()
    for d_eq, d_cl in outs:
        ctx.equalize_hist_device(d_in, d_eq, n, pitch, W, H, stream=sa)
        ctx.clahe_device(d_in, d_cl, n, pitch, W, H, 2.0, (8, 8), stream=sb)
        ctx.equalize_hist_device(d_in, d_eq, n, pitch, W, H, stream=sb)
        ctx.clahe_device(d_in, d_cl, n, pitch, W, H, 2.0, (8, 8), stream=sa)
    torch.cuda.synchronize()
    want_eq = oracle.c_nv12_batch("equalize", frames, W, H)
    want_cl = oracle.c_nv12_batch("clahe", frames, W, H, clip=2.0, tx=8, ty=8)
    for d_eq, d_cl in outs:
        assert np.array_equal(d_eq.cpu().numpy(), want_eq)
        assert np.array_equal(d_cl.cpu().numpy(), want_cl)
