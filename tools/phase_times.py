#!/usr/bin/env python
"""Time the phases of equalizeHist in isolation on device-resident 4K frames (GPU box only): histogram only,
LUT+apply only (external histogram), fused, and a plain device copy of the same bytes for reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import opencv_opencl_b200 as nv12eq

W, H, n = 3840, 2160, 256
pitch = nv12eq.nv12_frame_bytes(W, H)
ctx = nv12eq.Context(0, W, H, 1)
st = torch.cuda.current_stream()
d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
hist = torch.zeros(n * 256, dtype=torch.int32, device="cuda")


def timeit(name, fn, bytes_moved, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:42s} {ms:7.3f} ms/batch {ms * 1e3 / n:6.2f} us/frame  {bytes_moved / (ms * 1e-3) / 1e9:7.0f} GB/s moved", flush=True)


for ctas in (2, 3):
    ctx.set_tuning(0, 0, ctas, 0)
    print(f"--- ctas/SM = {ctas}")
    timeit("hist only (Y read)", lambda: (hist.zero_(), ctx.hist_device(d_in, n, pitch, W, H, hist, stream=st)), n * W * H)
    ctx.hist_device(d_in, n, pitch, W, H, hist.zero_(), stream=st)
    timeit("apply only (Y read + Y write, ext hist)", lambda: ctx.equalize_apply_device(d_in, d_out, n, pitch, W, H, hist, W * H, stream=st), 2 * n * W * H)
    timeit("fused equalizeHist (NV12 in, NV12 out)", lambda: ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st), 2 * n * pitch)
    timeit("fused, uv skipped", lambda: ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, uv_mode=2, stream=st), 2 * n * W * H)
timeit("torch copy of the NV12 batch", lambda: d_out.copy_(d_in), 2 * n * pitch)
y = d_in[: n * W * H]
timeit("torch copy, Y bytes only", lambda: d_out[: n * W * H].copy_(y), 2 * n * W * H)
timeit("torch read-only (sum of int32 view)", lambda: d_in.view(torch.int32).sum(), n * pitch)
