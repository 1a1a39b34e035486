#!/usr/bin/env python
"""Small fixed workload for ncu: a few launches of one op on device-resident synthetic frames (GPU box only)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import opencv_opencl_b200 as nv12eq  # noqa: E402

SIZES = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}
ap = argparse.ArgumentParser()
ap.add_argument("--op", default="equalize")
ap.add_argument("--size", default="4k")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--chunks", type=int, default=0)
ap.add_argument("--lag", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0)
ap.add_argument("--schedule", type=int, default=0)
a = ap.parse_args()
W, H = SIZES[a.size]
n, pitch = a.frames, (3 * W * H if a.op == "color" else nv12eq.nv12_frame_bytes(W, H))
ctx = nv12eq.Context(0, W, H, 1)
ctx.set_tuning(a.chunks, a.lag, a.ctas, a.schedule)
st = torch.cuda.current_stream()
d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
if a.op == "color":
    ctx.synth_bgr_device(d_in, n, pitch, W, H, stream=st)
else:
    ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
for _ in range(a.launches):
    if a.op == "equalize":
        ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
    elif a.op == "color":
        ctx.color_equalize_device(d_in, d_out, n, pitch, W, H, stream=st)
    else:
        ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (8, 8), stream=st)
torch.cuda.synchronize()
print("done", a.op, a.size, n)
