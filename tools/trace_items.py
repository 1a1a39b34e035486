#!/usr/bin/env python
"""Developer tool: run one CLAHE batch with NV12EQ_TRACE set and print a per-item-kind timing summary.
Needs a library built with the item trace compiled in:  make -C opencv-opencl_b200/csrc OUT=../libnv12eq_trace.so EXTRA=-DNV12EQ_ITEM_TRACE=1
and NV12EQ_LIB pointing at it."""
import os, sys, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = "/tmp/nv12eq_trace.bin"
os.environ["NV12EQ_TRACE"] = path
import torch
import opencv_opencl_b200 as nv12eq
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}[sys.argv[3] if len(sys.argv) > 3 else "4k"]
lag = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pitch = nv12eq.nv12_frame_bytes(W, H)
ctx = nv12eq.Context(0, W, H, 1)
ctx.set_tuning(0, lag, 0, 0)
st = torch.cuda.current_stream()
d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
for _ in range(2):
    ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (8, 8), stream=st)
torch.cuda.synchronize()
raw = open(path, "rb").read()
hdr = struct.unpack("8q", raw[:64]); items, per_slot, T, I, U, lag, grid, n = hdr
a = np.frombuffer(raw[64:], dtype=np.uint64).reshape(items, 4).astype(np.int64)
kind = a[:, 3] & 0xff; smid = a[:, 3] >> 8
valid = kind > 0
t0 = a[valid, 0].min()
print(f"items={items} per_slot={per_slot} T={T} I={I} U={U} lag={lag} grid={grid} frames={n}; total {(a[valid,2].max()-t0)/1e3:.1f} us -> {(a[valid,2].max()-t0)/1e3/n:.2f} us/frame")
for k, name in ((1, "tile"), (2, "cell"), (3, "uv")):
    m = (kind == k) & (a[:, 2] > 0)
    if not m.any(): continue
    dur = (a[m, 2] - a[m, 0]) / 1e3; wait = (a[m, 1] - a[m, 0]) / 1e3
    print(f"{name:5s} n={m.sum():6d} duration us: mean {dur.mean():7.2f} p50 {np.median(dur):7.2f} p90 {np.percentile(dur,90):7.2f} max {dur.max():7.2f} | wait mean {wait.mean():6.2f} p90 {np.percentile(wait,90):6.2f} max {wait.max():6.2f} | sum {dur.sum()/1e3:8.2f} ms")
m = valid & (a[:, 2] > 0)
busy = (a[m, 2] - a[m, 0]).sum() / 1e3
print(f"CTA-busy time {busy/1e3:.2f} ms over {grid} CTAs = {busy/grid:.1f} us per CTA; span {(a[m,2].max()-t0)/1e3:.1f} us")
# per-frame: when did tiles finish vs cells start
slot = np.arange(items) // per_slot; r = np.arange(items) % per_slot
for g in (5, 20, 40):
    if g + lag >= n: continue
    tm = (slot == g) & (r < T)
    cm = (slot == g + lag) & (r >= T) & (r < T + I)
    print(f"frame {g}: tiles start {(a[tm,0].min()-t0)/1e3:8.1f} .. end {(a[tm,2].max()-t0)/1e3:8.1f} us | cells start {(a[cm,0].min()-t0)/1e3:8.1f} first-ready {(a[cm,1].min()-t0)/1e3:8.1f} end {(a[cm,2].max()-t0)/1e3:8.1f}")
