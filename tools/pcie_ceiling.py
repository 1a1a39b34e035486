#!/usr/bin/env python
"""What the box's PCIe link gives with pinned memory (GPU box only): H2D alone, D2H alone, and both at once -- the ceiling
of the e2e leg of bench.py (which moves W*H bytes per frame each way)."""
import time
import torch
n = 2 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9


run(True, True, 1)
print(f"H2D alone {run(True, False):.1f} GB/s | D2H alone {run(False, True):.1f} GB/s | both at once {run(True, True):.1f} GB/s per direction")
