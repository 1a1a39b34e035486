"""CPU tests of the multi-GPU host logic: partitions, in-order reassembly, the round-robin
stream dispatcher, and -- under torch.distributed with the gloo backend, world size 2 -- the frame-sharded batch and
the spatially split single-frame mode with its 256-bin histogram all-reduce.  The arithmetic in the multi-process
tests is done by the oracle (test infrastructure standing in for the GPU); what is under test is the partitioning and
the exchange."""
import os
import random
import socket

import numpy as np
import pytest

import opencv_opencl_b200 as nv12eq

sh = nv12eq.sharding


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 2, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_stream_owner_and_frames_of_rank():
    world = 4
    seen = sorted(k for r in range(world) for k in sh.frames_of_rank(37, r, world))
    assert seen == list(range(37))
    assert all(sh.stream_owner(k, world) == r for r in range(world) for k in sh.frames_of_rank(37, r, world))


def test_row_bands_cover_the_plane_on_even_rows():
    for h in (2, 16, 1078, 1080, 2160, 1079):
        for world in (1, 2, 3, 8):
            bands = sh.row_bands(h, world, 2)
            assert len(bands) == world
            row = 0
            for first, rows in bands:
                assert first == row and rows >= 0
                row += rows
            assert row == h
            assert all(first % 2 == 0 for first, _ in bands)


def test_reassembler_restores_capture_order():
    rng = random.Random(7)
    for trial in range(20):
        n = 200
        order = list(range(n))
        # bounded disorder, like N GPUs finishing out of step
        for i in range(0, n, 8):
            window = order[i:i + 8]
            rng.shuffle(window)
            order[i:i + 8] = window
        ra = sh.Reassembler(max_reorder=16)
        out = []
        for seq in order:
            out += ra.push(seq, f"f{seq}")
        out += ra.flush()
        assert [s for s, _ in out] == list(range(n))
        assert all(item == f"f{s}" for s, item in out)
        assert ra.delivered == n and ra.dropped_late == 0 and ra.skipped == 0 and ra.max_held <= 16


def test_reassembler_lost_late_and_dropped_frames():
    ra = sh.Reassembler(max_reorder=3)
    out = []
    for seq in (0, 2, 3, 4):       # 1 is missing
        out += ra.push(seq, seq)
    assert [s for s, _ in out] == [0] and ra.held == 3
    out += ra.push(5, 5)           # 4 results wait behind the gap > max_reorder: give up on 1
    assert [s for s, _ in out] == [0, 2, 3, 4, 5] and ra.skipped == 1
    assert ra.push(1, 1) == [] and ra.dropped_late == 1       # arrives too late
    assert ra.push(5, 5) == [] and ra.dropped_late == 2       # duplicate
    assert ra.mark_dropped(6) == []                           # producer dropped 6 under back-pressure
    assert [s for s, _ in ra.push(7, 7)] == [7]
    assert ra.skipped == 2
    assert ra.flush() == []


class _FakeStream:
    """FIFO stand-in for nv12eq.Stream with a bounded depth and the drop-newest policy."""

    def __init__(self, depth, tag):
        self.depth, self.tag, self.q, self.seq, self.closed = depth, tag, [], 0, False

    def push(self, frame):
        s = self.seq
        self.seq += 1
        if len(self.q) >= self.depth:
            return None
        self.q.append((s, (self.tag, frame)))
        return s

    def pop(self, out=None, block=True):
        return self.q.pop(0) if self.q else None

    def close(self):
        self.closed = True


def test_frame_sharded_stream_round_robin_in_order_with_drops():
    streams = [_FakeStream(2, g) for g in range(3)]
    fs = sh.FrameShardedStream(streams)
    accepted = [fs.push(k) for k in range(9)]          # 3 GPUs x depth 2: frames 6, 7, 8 are dropped
    assert accepted == [0, 1, 2, 3, 4, 5, None, None, None] and fs.dropped == 3
    got = []
    while True:
        r = fs.pop()
        if r is None:
            break
        got.append(r)
    assert [k for k, _ in got] == [0, 1, 2, 3, 4, 5]
    assert [item for _, item in got] == [(k % 3, k) for k in range(6)]   # frame k ran on GPU k mod 3
    assert fs.pending() == 0
    assert fs.push(9) == 9 and fs.pop() == (9, (0, 9))
    fs.close()
    assert all(s.closed for s in streams)


# ------------------------------------------------------------------------------------------------------------
# torch.distributed / gloo, world size 2
# ------------------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, n = 96, 64, 7
        frames = np.stack([O.c_synth_nv12(W, H, 2026, k) for k in range(n)])

        # (1) frame-sharded batch: contiguous slices, no collective on the data path; gather only to check the result
        start, count = sh.shard_range(n, rank, world)
        mine = [O.c_nv12_equalize_hist(frames[k], W, H) for k in range(start, start + count)]
        gathered = [None] * world
        dist.all_gather_object(gathered, (start, mine))
        whole = np.stack([f for _, part in sorted(gathered, key=lambda t: t[0]) for f in part])
        want = np.stack([O.c_nv12_equalize_hist(frames[k], W, H) for k in range(n)])
        ok_batch = bool(np.array_equal(whole, want))

        # (2) stream mode: frame k on rank k mod world, results reach rank 0 out of order, Reassembler restores order
        done = [(k, O.c_nv12_clahe(frames[k], W, H, 2.0, 4, 4)) for k in sh.frames_of_rank(n, rank, world)]
        lists = [None] * world
        dist.all_gather_object(lists, done)
        arrivals = [x for r in reversed(range(world)) for x in lists[r]]  # rank 1's results arrive first
        ra = sh.Reassembler(max_reorder=8)
        ordered = []
        for k, f in arrivals:
            ordered += ra.push(k, f)
        ordered += ra.flush()
        ok_stream = [k for k, _ in ordered] == list(range(n)) and all(
            np.array_equal(f, O.c_nv12_clahe(frames[k], W, H, 2.0, 4, 4)) for k, f in ordered)

        # (3) spatial split of ONE frame: band histograms -> all-reduce (the only collective) -> same LUT everywhere
        y = frames[0][:W * H].reshape(H, W)
        first, rows = sh.row_bands(H, world, 2)[rank]
        band = y[first:first + rows]
        hist = torch.from_numpy(O.c_hist256(band).astype(np.int32))
        sh.allreduce_histograms(hist)
        ok_hist = bool(np.array_equal(hist.numpy(), O.c_hist256(y)))
        lut = O.c_equalize_lut(hist.numpy(), W * H)
        out_band = lut[band]
        bands = [None] * world
        dist.all_gather_object(bands, (first, out_band))
        full = np.concatenate([b for _, b in sorted(bands, key=lambda t: t[0])])
        ok_split = bool(np.array_equal(full, O.c_equalize_hist(y)))

        # (4) spatial split of ONE frame for CLAHE: band tile LUTs -> one LUT row to each neighbour (send/recv) -> band interpolation
        tx, ty, clip = 4, 4, 2.0
        th = H // ty
        first_t, nb = sh.tile_row_bands(ty, world)[rank]
        band = y[first_t * th:(first_t + nb) * th]
        halo = torch.full(((nb + 2) * tx * 256,), 0xEE, dtype=torch.uint8)
        halo[tx * 256:(nb + 1) * tx * 256] = torch.from_numpy(O.c_clahe_tile_luts(band, clip, tx, nb).reshape(-1))
        sh.exchange_lut_halo(halo, tx, nb, rank, world)
        out_band = O.c_clahe_interp_band(band, H, tx, ty, first_t * th, halo.numpy(), first_t)
        bands = [None] * world
        dist.all_gather_object(bands, (first_t, out_band))
        full = np.concatenate([b for _, b in sorted(bands, key=lambda t: t[0])])
        ok_clahe = bool(np.array_equal(full, O.c_clahe(y, clip, tx, ty)))

        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as f:
            f.write(f"{ok_batch} {ok_stream} {ok_hist} {ok_split} {ok_clahe}")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_frame_sharding_and_spatial_split(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "True True True True True", f"rank {r}"
