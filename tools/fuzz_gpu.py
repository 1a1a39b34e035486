#!/usr/bin/env python
"""Soak test (GPU box only): random geometries, batch sizes, strides, uv modes and tuning knobs through the device-resident and
the host-buffer entry points for a time budget, every result compared bit-exactly with the oracle.
   python tools/fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import opencv_opencl_b200 as nv  # noqa: E402
from oracle import oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
t0 = time.time()
cases = fails = 0
st = torch.cuda.current_stream()
with nv.Context(0, 4096, 2304, 3) as ctx:
    while time.time() - t0 < budget:
        cases += 1
        W = int(rng.choice([16, 32, 48, 64, 96, 120, 128, 160, 240, 256, 320, 322, 480, 512, 640, 642, 960, 1280, 1920]))
        H = int(rng.choice([8, 16, 18, 32, 36, 64, 90, 128, 135, 180, 270, 360, 540, 720, 1080]))
        tx, ty = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 12, 16])), int(rng.choice([1, 2, 3, 4, 6, 8, 9, 16]))
        clip = float(rng.choice([0.0, 0.7, 2.0, 3.0, 40.0]))
        n = int(rng.integers(1, 7))
        S = W + int(rng.choice([0, 0, 0, 16, 32, 8, 3]))
        uv = int(rng.integers(0, 3))
        pitch = S * (H + H // 2) + int(rng.choice([0, 0, 64, 16]))
        ctx.set_tuning(0, int(rng.choice([0, 0, -1, 1, 2, 5])), int(rng.choice([0, 0, 1, 2])), 0)
        frames = np.zeros(n * pitch, np.uint8)
        for k in range(n):
            frames[k * pitch:k * pitch + S * (H + H // 2)] = O.c_synth_nv12(W, H, int(rng.integers(1, 1 << 20)), k, stride=S)
        d_in = torch.from_numpy(frames).cuda()
        pre = rng.integers(0, 256, frames.size, dtype=np.uint8)
        for op in ("clahe", "equalize"):
            d_out = torch.from_numpy(pre.copy()).cuda()
            if op == "clahe":
                ctx.clahe_device(d_in, d_out, n, pitch, W, H, clip, (tx, ty), stride=S, uv_mode=uv, stream=st)
            else:
                ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stride=S, uv_mode=uv, stream=st)
            got = d_out.cpu().numpy()
            for k in range(n):
                a, b = k * pitch, k * pitch + S * (H + H // 2)
                if op == "clahe":
                    want = O.c_nv12_clahe(frames[a:b], W, H, clip, tx, ty, stride=S, uv_mode=uv, out=pre[a:b].copy())
                else:
                    want = O.c_nv12_equalize_hist(frames[a:b], W, H, stride=S, uv_mode=uv, out=pre[a:b].copy())
                if not np.array_equal(got[a:b], want):
                    fails += 1
                    print("MISMATCH", op, dict(W=W, H=H, S=S, tx=tx, ty=ty, clip=clip, n=n, uv=uv, pitch=pitch, frame=k), flush=True)
            if not np.array_equal(got.reshape(-1)[[i for k in range(n) for i in range(k * pitch + S * (H + H // 2), (k + 1) * pitch)]],
                                  pre.reshape(-1)[[i for k in range(n) for i in range(k * pitch + S * (H + H // 2), (k + 1) * pitch)]]):
                fails += 1
                print("GAP BYTES TOUCHED", op, dict(W=W, H=H, S=S, n=n, pitch=pitch), flush=True)
        if cases % 5 == 0 and W % 2 == 0 and H % 2 == 0:   # colour path and adapters on the same geometry
            bgr = O.c_synth_bgr(W, H, int(rng.integers(0, 50)))
            mode = int(rng.integers(0, 2))
            checks = [("color_equalize", ctx.color_equalize(bgr, mode), O.c_color_equalize(bgr, mode)),
                      ("color_clahe", ctx.color_clahe(bgr, clip, (tx, ty), mode), O.c_color_equalize(bgr, mode, True, clip, tx, ty)),
                      ("bgr_to_i420", ctx.bgr_to_i420(bgr), O.c_bgr2i420(bgr)),
                      ("bgr_to_nv12", ctx.bgr_to_nv12(bgr), O.c_bgr_to_nv12(bgr))]
            fr0 = np.ascontiguousarray(frames[:S * (H + H // 2)].reshape(-1, S)[:, :W]).reshape(-1)
            checks.append(("nv12_to_bgr", ctx.nv12_to_bgr(fr0, W, H), O.c_nv12_to_bgr(fr0, W, H)))
            for name, got_, want_ in checks:
                if not np.array_equal(got_, want_):
                    fails += 1
                    print("MISMATCH", name, dict(W=W, H=H, mode=mode, clip=clip, tx=tx, ty=ty), flush=True)
        if cases % 7 == 0:   # host-buffer form of the same frame
            fr = frames[:S * (H + H // 2)].copy()
            got = ctx.clahe(fr, W, H, clip, (tx, ty), stride=S, uv_mode=uv, out=pre[:fr.size].copy())
            want = O.c_nv12_clahe(fr, W, H, clip, tx, ty, stride=S, uv_mode=uv, out=pre[:fr.size].copy())
            if not np.array_equal(got, want):
                fails += 1
                print("MISMATCH host clahe", dict(W=W, H=H, S=S, tx=tx, ty=ty, clip=clip, uv=uv), flush=True)
print(f"fuzz: {cases} cases in {time.time() - t0:.0f} s, {fails} failures (seed {seed})")
sys.exit(1 if fails else 0)
