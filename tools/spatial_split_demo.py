#!/usr/bin/env python
"""Optional spatially split single-frame mode on real GPUs under torchrun / NCCL:
every rank holds a band of rows of ONE frame, histograms it (nv12eq_hist_device), the 256-bin histograms are summed
with one NCCL all-reduce (the only collective anywhere on this path), and every rank applies the LUT of the summed
histogram to its band (nv12eq_equalize_apply_device).  Rank 0 gathers the bands, checks them bit-exactly against the
oracle and prints one JSON line with the time per frame (CUDA events, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/spatial_split_demo.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import opencv_opencl_b200 as nv12eq  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nv12eq.build()
    from oracle import oracle as O
    W, H = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))   # an 8K luma plane: 33 MB
    ctx = nv12eq.Context(local, W, H, 1)
    st = torch.cuda.current_stream()
    y = O.c_synth_nv12(W, H, 2026, 0)[:W * H].reshape(H, W)       # every rank synthesises the frame, keeps its band
    first, rows = nv12eq.sharding.row_bands(H, world, 2)[rank]
    d_in = torch.from_numpy(np.ascontiguousarray(y[first:first + rows])).cuda().reshape(-1)
    d_out = torch.zeros_like(d_in)
    eq = nv12eq.sharding.SpatialSplitEqualizer(ctx, W, H, rank, world)
    for _ in range(3):
        eq.run(d_in, d_out, stream=st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record(st)
    for _ in range(iters):
        eq.run(d_in, d_out, stream=st)
    e1.record(st)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    band = d_out.cpu().numpy().reshape(rows, W)
    bands = [None] * world
    if world > 1:
        dist.all_gather_object(bands, (first, band))
    else:
        bands = [(first, band)]
    if rank == 0:
        full = np.concatenate([b for _, b in sorted(bands, key=lambda t: t[0])])
        ok = bool(np.array_equal(full, O.c_equalize_hist(y)))
        print(json.dumps({"mode": "spatial split of one frame, 256-bin histogram all-reduce (NCCL)", "n_gpus": world,
                          "frame": f"{W}x{H} luma plane", "bands": [list(nv12eq.sharding.row_bands(H, world, 2)[r]) for r in range(world)],
                          "ms_per_frame": float(ms.item()), "bit_exact_vs_oracle": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
