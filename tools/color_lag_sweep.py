#!/usr/bin/env python
"""Fused colour equalization: time per frame vs the frame lag between the histogram pass and the apply pass and the number
of chunks per frame (GPU box only).  Usage: color_lag_sweep.py [lags] [chunks] [4k|1080p]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, opencv_opencl_b200 as nv
W, H = {"4k": (3840, 2160), "1080p": (1920, 1080)}[sys.argv[3] if len(sys.argv) > 3 else "4k"]
n = 128 if W > 2000 else 256
pitch = 3 * W * H
c = nv.Context(0, W, H, 1); st = torch.cuda.current_stream()
a = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
c.synth_bgr_device(a, n, pitch, W, H, stream=st)
lags = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4,6").split(",")]
chunk_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
for chunks, lag in [(ch, lg) for ch in chunk_list for lg in lags]:
    time.sleep(0.5)   # let the board cool between configurations: later ones would otherwise run power-capped
    c.set_tuning(chunks, lag, 0, 0)
    for _ in range(3): c.color_equalize_device(a, b, n, pitch, W, H, stream=st)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): c.color_equalize_device(a, b, n, pitch, W, H, stream=st)
    e1.record(st); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    print(f"chunks {chunks} lag {lag}: {ms / n * 1e3:.2f} us/frame {n / ms * 1e3:.0f} fps frac {n * 6 * W * H / ms / 1e6 / 6547.2:.3f}")
