#!/usr/bin/env python
"""Device-side latency of small batches (1, 2, 4, 8 frames per launch): what a 60 fps stream sees per frame (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opencv_opencl_b200 as nv
st = torch.cuda.current_stream()
for (W, H, name) in ((3840, 2160, "4k"), (1920, 1080, "1080p")):
    pitch = nv.nv12_frame_bytes(W, H)
    c = nv.Context(0, W, H, 1)
    a = torch.empty(8 * pitch, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    c.synth_nv12_device(a, 8, pitch, W, H, stream=st)
    for op in ("equalize", "clahe"):
        for n in (1, 2, 4, 8):
            f = (lambda: c.equalize_hist_device(a, b, n, pitch, W, H, stream=st)) if op == "equalize" else \
                (lambda: c.clahe_device(a, b, n, pitch, W, H, 2.0, (8, 8), stream=st))
            for _ in range(5): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(50): f()
            e1.record(st); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 50 * 1e3
            print(f"{op:8s} {name:5s} n={n}: {us:8.1f} us per launch, {us / n:7.1f} us per frame")
    c.close()
