"""Multi-GPU partitioning of the NV12 hot path.  Host logic only -- no arithmetic here.

Frames are independent units (the reference already runs frame-level data parallelism over worker threads,
OpenCVequalHist.cpp:397-402), so the N-GPU path is a partition, not a collective:

* batch mode   -- rank r of N owns a contiguous, balanced slice of the frame index (`shard_range`);
* stream mode  -- frame k goes to GPU k mod N (`stream_owner`) and finished frames are put back in capture order by a
                  sequence-numbered reorder buffer (`Reassembler`), which is what the reference's unpublished
                  IMP/improvement binaries added ("frame-output-ordering", "Max reorder", "Dropped: late");
                  `FrameShardedStream` does this in one process over one `nv12eq` stream per GPU;
* spatial split of ONE frame (optional, latency bound) -- every rank histograms its band of rows, the 256-bin
  histograms are summed with one all-reduce (the only collective anywhere on this path), every rank applies the LUT of
  the summed histogram to its band (`row_bands`, `allreduce_histograms`, `SpatialSplitEqualizer`).

`torch.distributed` is plumbing: NCCL on the GPU box, gloo in the CPU tests (tests/test_sharding.py, world size 2).
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple


# ------------------------------------------------------------------------------------------------------------
# partitions
# ------------------------------------------------------------------------------------------------------------
def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: returns (start, count); the first n_items % world ranks get one item more."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad partition request n={n_items} rank={rank} world={world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def stream_owner(seq: int, world: int) -> int:
    """Frame `seq` of a stream is processed by GPU seq mod world."""
    return seq % world


def frames_of_rank(n_frames: int, rank: int, world: int) -> range:
    """Stream mode: the sequence numbers rank `rank` processes."""
    return range(rank, n_frames, world)


def row_bands(height: int, world: int, align: int = 2) -> List[Tuple[int, int]]:
    """Spatial split of one plane into `world` bands of rows: [(first_row, n_rows)], boundaries on multiples of
    `align` (2 keeps NV12 chroma rows with their luma rows).  Bands may be empty when height < world*align."""
    units = height // align
    bands = []
    for r in range(world):
        s, c = shard_range(units, r, world)
        bands.append((s * align, c * align))
    last_start, last_rows = bands[-1]
    bands[-1] = (last_start, height - last_start)  # the remainder rows (height % align) go to the last band
    return bands


# ------------------------------------------------------------------------------------------------------------
# in-order reassembly
# ------------------------------------------------------------------------------------------------------------
class Reassembler:
    """Sequence-numbered reorder buffer: results arrive in any order, come out in capture order.

    max_reorder bounds the buffer: when more than max_reorder results are waiting behind a missing sequence number,
    the missing frame is declared lost (`skipped`) and delivery moves on -- a stalled GPU cannot stall the stream.
    A result whose number has already been passed is dropped as late (`dropped_late`), as in the reference's IMP binary.
    """

    def __init__(self, max_reorder: int = 16, first_seq: int = 0):
        if max_reorder < 1:
            raise ValueError("max_reorder must be >= 1")
        self.max_reorder = max_reorder
        self.next_seq = first_seq
        self._held: Dict[int, Any] = {}
        self.delivered = 0
        self.dropped_late = 0
        self.skipped = 0
        self.max_held = 0

    def push(self, seq: int, item: Any) -> List[Tuple[int, Any]]:
        """Hand in one finished result; returns the (seq, item) pairs that became deliverable, in order."""
        if seq < self.next_seq or seq in self._held:
            self.dropped_late += 1
            return []
        self._held[seq] = item
        self.max_held = max(self.max_held, len(self._held))
        out = self._drain()
        while len(self._held) > self.max_reorder:
            # give up on the oldest missing frame(s): jump to the smallest held number
            nxt = min(self._held)
            self.skipped += nxt - self.next_seq
            self.next_seq = nxt
            out += self._drain()
        return out

    def mark_dropped(self, seq: int) -> List[Tuple[int, Any]]:
        """The producer knows frame `seq` will never arrive (back-pressure drop): do not wait for it."""
        if seq >= self.next_seq and seq not in self._held:
            self._held[seq] = _DROPPED
        return self._drain()

    def flush(self) -> List[Tuple[int, Any]]:
        """End of stream: deliver whatever is held, in order, skipping the gaps."""
        out = []
        for seq in sorted(self._held):
            item = self._held.pop(seq)
            self.skipped += seq - self.next_seq
            self.next_seq = seq + 1
            if item is not _DROPPED:
                out.append((seq, item))
                self.delivered += 1
            else:
                self.skipped += 1
        return out

    def _drain(self) -> List[Tuple[int, Any]]:
        out = []
        while self.next_seq in self._held:
            item = self._held.pop(self.next_seq)
            if item is _DROPPED:
                self.skipped += 1
            else:
                out.append((self.next_seq, item))
                self.delivered += 1
            self.next_seq += 1
        return out

    @property
    def held(self) -> int:
        return len(self._held)


class _Dropped:
    def __repr__(self):
        return "<dropped>"


_DROPPED = _Dropped()


# ------------------------------------------------------------------------------------------------------------
# one process, several GPUs: frame k -> GPU k mod N, in-order pop
# ------------------------------------------------------------------------------------------------------------
class FrameShardedStream:
    """Round-robin a frame stream over several per-GPU streams and deliver in capture order.

    `streams` are objects with push(frame) -> seq | None, pop(out=None, block=True) -> (seq, frame) | None and close()
    (``nv12eq.Stream`` instances, one per GPU context).  Every per-GPU stream is FIFO, so popping stream (k mod N) for
    k = 0, 1, 2, ... yields capture order without a reorder buffer; frames dropped by a stream's back-pressure policy
    show up as gaps and are skipped.
    """

    def __init__(self, streams: Sequence[Any]):
        if not streams:
            raise ValueError("need at least one stream")
        import threading
        self.streams = list(streams)
        self.world = len(self.streams)
        self._push_seq = 0
        self._pop_seq = 0
        self._local: Dict[int, Optional[int]] = {}   # global seq -> per-stream seq (None = dropped)
        self.dropped = 0
        self._cv = threading.Condition()              # one producer thread may push while one consumer thread pops

    def push(self, frame) -> Optional[int]:
        k = self._push_seq
        local = self.streams[stream_owner(k, self.world)].push(frame)   # may block (FULL_BLOCK): outside the lock
        with self._cv:
            self._local[k] = local
            self._push_seq = k + 1
            if local is None:
                self.dropped += 1
            self._cv.notify_all()
        return None if local is None else k

    def pop(self, out=None, block: bool = True, wait_push: bool = False):
        """Next frame in capture order: (global_seq, frame).  None when nothing has been pushed that is not yet popped
        (unless wait_push: then wait for the producer thread), or when block=False and the frame is not finished."""
        while True:
            with self._cv:
                while self._pop_seq >= self._push_seq:
                    if not wait_push:
                        return None
                    self._cv.wait()
                k = self._pop_seq
                local = self._local.pop(k)
                if local is None:          # dropped at push time: skip the gap
                    self._pop_seq += 1
                    continue
            got = self.streams[stream_owner(k, self.world)].pop(out=out, block=block)
            if got is None:                # not finished yet (block=False): put the bookkeeping back
                with self._cv:
                    self._local[k] = local
                return None
            seq, frame = got
            if seq > local:                # the owner discarded frame k (drop-oldest) and delivered a later one
                raise RuntimeError("per-GPU stream skipped ahead; use FULL_BLOCK or FULL_DROP_NEWEST with FrameShardedStream")
            with self._cv:
                self._pop_seq += 1
            return k, frame

    def pending(self) -> int:
        return self._push_seq - self._pop_seq

    def close(self):
        for s in self.streams:
            s.close()


# ------------------------------------------------------------------------------------------------------------
# one process, several GPUs: a host batch cut into contiguous slices, one slice per GPU, no collective
# ------------------------------------------------------------------------------------------------------------
class ShardedBatch:
    """Frame-sharded batch processing over several contexts (one per GPU) from ONE process: rank r of N gets the
    contiguous slice `shard_range(n_frames, r, N)` of the host batch and runs it through its own context on its own host
    thread (the C calls release the GIL), so uploads, kernels and downloads of all GPUs overlap.  The N-process form of
    the same partition is what bench.py runs under torchrun."""

    def __init__(self, contexts: Sequence[Any]):
        if not contexts:
            raise ValueError("need at least one context")
        from concurrent.futures import ThreadPoolExecutor
        self.contexts = list(contexts)
        self.world = len(self.contexts)
        self._pool = ThreadPoolExecutor(max_workers=self.world)

    def _run(self, method: str, frames, out, n_frames: int, frame_pitch: int, width: int, height: int, **kw):
        import numpy as np
        flat_in = np.asarray(frames).reshape(-1)
        flat_out = np.asarray(out).reshape(-1)

        def one(rank):
            start, count = shard_range(n_frames, rank, self.world)
            if count == 0:
                return
            a, b = start * frame_pitch, (start + count) * frame_pitch
            getattr(self.contexts[rank], method)(flat_in[a:b], width, height, out=flat_out[a:b], n_frames=count,
                                                 frame_pitch=frame_pitch, **kw)
        list(self._pool.map(one, range(self.world)))   # re-raises the first worker exception
        return out

    def equalize_hist(self, frames, width, height, out, n_frames, frame_pitch, **kw):
        return self._run("equalize_hist_batch", frames, out, n_frames, frame_pitch, width, height, **kw)

    def clahe(self, frames, width, height, out, n_frames, frame_pitch, clip_limit=2.0, tiles=(8, 8), **kw):
        return self._run("clahe_batch", frames, out, n_frames, frame_pitch, width, height, clip_limit=clip_limit, tiles=tiles, **kw)

    def close(self):
        self._pool.shutdown(wait=True)


# ------------------------------------------------------------------------------------------------------------
# optional spatial split of one frame: the only collective on the path
# ------------------------------------------------------------------------------------------------------------
def allreduce_histograms(hist, group=None):
    """Sum 256-bin histograms (a torch tensor [..., 256], int32/int64, on the backend's device) over all ranks in place.
    1 KB per frame: pure latency (NVLink/NVSwitch all-reduce ~10-20 us vs ~5 us of work per 4K frame)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


class SpatialSplitEqualizer:
    """equalizeHist of ONE frame split by rows over the ranks of a process group (GPU path; needs libnv12eq + CUDA).

    Every rank holds its band of the Y plane on its GPU.  `run` = nv12eq_hist_device on the band -> all-reduce of the
    256-bin histogram -> nv12eq_equalize_apply_device on the band with the whole frame's pixel count.
    """

    def __init__(self, ctx, width: int, height: int, rank: int, world: int, group=None):
        self.ctx, self.width, self.height, self.rank, self.world, self.group = ctx, width, height, rank, world, group
        self.band = row_bands(height, world, 2)[rank]

    def run(self, d_band_in, d_band_out, stream=None):
        """`stream`: a torch.cuda.Stream, or None for torch's current stream.  Everything -- the zero fill of the histogram,
        the histogram kernel, the all-reduce and the apply kernel -- is queued on that ONE stream (made torch's current
        stream for the duration), so the steps are ordered by the stream and nothing has to be drained in between."""
        import torch
        if stream is not None and not isinstance(stream, torch.cuda.Stream):
            raise TypeError("SpatialSplitEqualizer.run: stream must be a torch.cuda.Stream or None (torch's current stream)")
        st = stream if stream is not None else torch.cuda.current_stream(d_band_in.device)
        first, rows = self.band
        with torch.cuda.stream(st):
            hist = torch.zeros(256, dtype=torch.int32, device=d_band_in.device)
            if rows > 0:
                self.ctx.hist_device(d_band_in, 1, self.width * rows, self.width, rows, hist, stream=st)
            allreduce_histograms(hist, self.group)   # NCCL work is ordered on torch's current stream = st
            if rows > 0:
                self.ctx.equalize_apply_device(d_band_in, d_band_out, 1, self.width * rows, self.width, rows, hist,
                                               self.width * self.height, stream=st)
        return hist


def tile_row_bands(tiles_y: int, world: int) -> List[Tuple[int, int]]:
    """Spatial split of a CLAHE tile grid: [(first_tile_row, n_tile_rows)] per rank, contiguous, as even as possible."""
    return [shard_range(tiles_y, r, world) for r in range(world)]


def exchange_lut_halo(luts_halo, tiles_x: int, band_tiles_y: int, rank: int, world: int, group=None, has_rows=None):
    """The one exchange of the spatially split CLAHE: every rank sends its first tile row of LUTs up and its last tile row down
    and receives its neighbours' rows into the halo rows of `luts_halo` (a torch tensor [(band_tiles_y + 2) * tiles_x * 256] of
    uint8 on the backend's device; rows 1 .. band_tiles_y are the rank's own).  tiles_x * 256 bytes per message (2 KB for an
    8 x 8 grid): pure latency, like the 256-bin all-reduce of the equalizeHist split.  `has_rows[r]` says whether rank r owns any
    tile row (ranks without rows neither send nor receive; their neighbours talk to the next rank that has rows)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world <= 1 or band_tiles_y == 0:
        return luts_halo
    has = list(has_rows) if has_rows is not None else [True] * world
    up = next((r for r in range(rank - 1, -1, -1) if has[r]), None)
    down = next((r for r in range(rank + 1, world) if has[r]), None)
    row = tiles_x * 256
    first, last = luts_halo[row:2 * row], luts_halo[band_tiles_y * row:(band_tiles_y + 1) * row]
    halo_top, halo_bot = luts_halo[0:row], luts_halo[(band_tiles_y + 1) * row:(band_tiles_y + 2) * row]
    ops = []
    if up is not None:
        ops += [dist.P2POp(dist.isend, first, up, group), dist.P2POp(dist.irecv, halo_top, up, group)]
    if down is not None:
        ops += [dist.P2POp(dist.isend, last, down, group), dist.P2POp(dist.irecv, halo_bot, down, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return luts_halo


class SpatialSplitClahe:
    """CLAHE of ONE frame split by tile rows over the ranks of a process group (GPU path; needs libnv12eq + CUDA).

    Every rank holds its band of the Y plane (whole tile rows; the tile grid must divide the frame) on its GPU.
    `run` = nv12eq_clahe_band_luts_device on the band -> exchange of one tile row of LUTs with each neighbour (NCCL send/recv,
    the only communication) -> nv12eq_clahe_band_apply_device.  Bit-exact with the single-GPU result: the tile LUTs depend on
    the tile's pixels only, and the interpolation uses the weights of the whole frame."""

    def __init__(self, ctx, width: int, height: int, clip_limit: float, tiles: Tuple[int, int], rank: int, world: int, group=None):
        if width % tiles[0] or height % tiles[1]:
            raise ValueError("the spatial split needs a tile grid that divides the frame")
        self.ctx, self.width, self.height, self.clip, self.tiles = ctx, width, height, clip_limit, tiles
        self.rank, self.world, self.group = rank, world, group
        self.bands = tile_row_bands(tiles[1], world)
        self.first_tile_row, self.band_tiles_y = self.bands[rank]
        th = height // tiles[1]
        self.band = (self.first_tile_row * th, self.band_tiles_y * th)   # (first row, rows) of this rank

    def run(self, d_band_in, d_band_out, stream=None):
        import torch
        if stream is not None and not isinstance(stream, torch.cuda.Stream):
            raise TypeError("SpatialSplitClahe.run: stream must be a torch.cuda.Stream or None (torch's current stream)")
        st = stream if stream is not None else torch.cuda.current_stream(d_band_in.device)
        tx, nb = self.tiles[0], self.band_tiles_y
        with torch.cuda.stream(st):
            luts = torch.zeros((nb + 2) * tx * 256, dtype=torch.uint8, device=d_band_in.device)
            if nb > 0:
                self.ctx.clahe_band_luts_device(d_band_in, self.width, self.height, self.clip, self.tiles, self.first_tile_row, nb,
                                                luts[tx * 256:], stream=st)
            exchange_lut_halo(luts, tx, nb, self.rank, self.world, self.group, [c > 0 for _, c in self.bands])
            if nb > 0:
                self.ctx.clahe_band_apply_device(d_band_in, d_band_out, self.width, self.height, self.tiles, self.first_tile_row, nb,
                                                 luts, stream=st)
        return luts
