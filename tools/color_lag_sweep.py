#!/usr/bin/env python
"""Fused colour equalization: time per frame vs the frame lag between the histogram pass and the apply pass (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, opencv_opencl_b200 as nv
W, H, n = 3840, 2160, 128
pitch = 3 * W * H
c = nv.Context(0, W, H, 1); st = torch.cuda.current_stream()
a = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
c.synth_bgr_device(a, n, pitch, W, H, stream=st)
for lag in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4,6").split(",")]:
    c.set_tuning(0, lag, 0, 0)
    for _ in range(3): c.color_equalize_device(a, b, n, pitch, W, H, stream=st)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): c.color_equalize_device(a, b, n, pitch, W, H, stream=st)
    e1.record(st); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    print(f"lag {lag}: {ms / n * 1e3:.2f} us/frame {n / ms * 1e3:.0f} fps frac {n * 6 * W * H / ms / 1e6 / 6547.2:.3f}")
