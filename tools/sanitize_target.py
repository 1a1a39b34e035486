#!/usr/bin/env python
"""Small workload for compute-sanitizer (GPU box only): every kernel of the library once or twice on small frames, results
checked against the oracle.  Usage:
   compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_target.py [--frames N]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import opencv_opencl_b200 as nv  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=6)
a = ap.parse_args()
n = a.frames
ok = True
st = torch.cuda.current_stream()
for (W, H, tiles) in ((320, 192, 4), (322, 200, 8)):   # the second size takes the CLAHE padding path and the unaligned row paths
    pitch = nv.nv12_frame_bytes(W, H)
    with nv.Context(0, W, H, 1) as ctx:
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros_like(d_in)
        ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
        ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
        torch.cuda.synchronize()
        for k in range(n):
            fr = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
            ok &= bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), O.c_nv12_equalize_hist(fr, W, H)))
        d_out.zero_()
        ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (tiles, tiles), stream=st)
        torch.cuda.synchronize()
        for k in range(n):
            fr = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
            ok &= bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), O.c_nv12_clahe(fr, W, H, 2.0, tiles, tiles)))
        # host-buffer forms (lanes, pinned staging)
        fr = O.c_synth_nv12(W, H, 2026, 3)
        ok &= bool(np.array_equal(ctx.equalize_hist(fr, W, H), O.c_nv12_equalize_hist(fr, W, H)))
        ok &= bool(np.array_equal(ctx.clahe(fr, W, H, 3.0, (tiles, tiles)), O.c_nv12_clahe(fr, W, H, 3.0, tiles, tiles)))
        # colour path, adapters
        bp = 3 * W * H
        d_bgr = torch.empty(n * bp, dtype=torch.uint8, device="cuda")
        d_bo = torch.zeros_like(d_bgr)
        ctx.synth_bgr_device(d_bgr, n, bp, W, H, stream=st)
        ctx.color_equalize_device(d_bgr, d_bo, n, bp, W, H, stream=st)
        torch.cuda.synchronize()
        for k in (0, n - 1):
            src = d_bgr[k * bp:(k + 1) * bp].cpu().numpy().reshape(H, W, 3)
            ok &= bool(np.array_equal(d_bo[k * bp:(k + 1) * bp].cpu().numpy().reshape(H, W, 3), O.c_color_equalize(src, O.COLOR_YUV)))
        bgr = O.c_synth_bgr(W, H, 1)
        ok &= bool(np.array_equal(ctx.bgr_to_i420(bgr), O.c_bgr2i420(bgr)))
        ok &= bool(np.array_equal(ctx.bgr_to_nv12(bgr), O.c_bgr_to_nv12(bgr)))
        ok &= bool(np.array_equal(ctx.nv12_to_bgr(fr, W, H), O.c_nv12_to_bgr(fr, W, H)))
        ok &= bool(np.array_equal(ctx.color_clahe(bgr, 2.0, (tiles, tiles)), O.c_color_equalize(bgr, O.COLOR_YUV, True, 2.0, tiles, tiles)))
print("sanitize_target:", "results match the oracle" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
