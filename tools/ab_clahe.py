#!/usr/bin/env python
"""A/B of library builds on the CLAHE (and equalizeHist / colour) device-resident workloads (GPU box only).

  python tools/ab_clahe.py --libs default,opencv-opencl_b200/libnv12eq_x.so --sizes 4k,1080p --lags 0 --ctas 0 --rounds 2

Every (build, size, lag, ctas) configuration runs in its own process (the library is chosen at import time through
NV12EQ_LIB), after a cool-down, `rounds` times in alternation -- back-to-back configurations drift as the board heats up.
Each run checks two frames of its output against the oracle (a fast wrong kernel is not a result) and prints
us per frame and the fraction of the measured HBM roofline."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}


def child(a):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import opencv_opencl_b200 as nv
    from oracle import oracle as O
    W, H = SIZES[a.size]
    n = a.frames
    color = a.op == "color"
    pitch = 3 * W * H if color else nv.nv12_frame_bytes(W, H)
    ctx = nv.Context(0, W, H, 1)
    ctx.set_tuning(a.chunks, a.lag, a.cta, 0)
    st = torch.cuda.current_stream()
    d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros_like(d_in)
    if color:
        ctx.synth_bgr_device(d_in, n, pitch, W, H, stream=st)
    else:
        ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)

    def step():
        if a.op == "clahe":
            ctx.clahe_device(d_in, d_out, n, pitch, W, H, a.clip, (a.tiles, a.tiles), stream=st)
        elif a.op == "equalize":
            ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
        else:
            ctx.color_equalize_device(d_in, d_out, n, pitch, W, H, stream=st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    e[0].record(st)
    for i in range(a.iters):
        step()
        e[i + 1].record(st)
    torch.cuda.synchronize()
    ms = sorted(e[i].elapsed_time(e[i + 1]) for i in range(a.iters))
    ok = True
    for k in (1, n - 1):
        fr = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
        want = (O.c_nv12_clahe(fr, W, H, a.clip, a.tiles, a.tiles) if a.op == "clahe" else O.c_nv12_equalize_hist(fr, W, H) if a.op == "equalize"
                else O.c_color_equalize(fr.reshape(H, W, 3), O.COLOR_YUV).reshape(-1))
        ok = ok and bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), want))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    mean = sum(ms) / len(ms)
    algo = (6 if color else 3) * W * H
    print(json.dumps({"lib": os.path.basename(os.environ.get("NV12EQ_LIB", "default")), "op": a.op, "size": a.size, "lag": a.lag, "ctas": a.cta,
                      "chunks": a.chunks, "tiles": a.tiles, "skip": os.environ.get("NV12EQ_DEBUG_SKIP", ""), "us_per_frame_mean": mean * 1e3 / n, "us_per_frame_min": ms[0] * 1e3 / n,
                      "frac_mean": n * algo / (mean * 1e-3) / 1e9 / peak, "parity": ok}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="default")
    ap.add_argument("--op", default="clahe")
    ap.add_argument("--sizes", default="4k")
    ap.add_argument("--lags", default="0")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--chunkss", default="0")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--tiles", type=int, default=8)
    ap.add_argument("--clip", type=float, default=2.0)
    ap.add_argument("--cooldown", type=float, default=2.0)
    ap.add_argument("--skips", default="", help="comma list of NV12EQ_DEBUG_SKIP masks (phase timing; parity is then expected to fail)")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--size", default="4k")
    ap.add_argument("--lag", type=int, default=0)
    ap.add_argument("--cta", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=0)
    a = ap.parse_args()
    if a.child:
        return child(a)
    for _ in range(a.rounds):
        for size in a.sizes.split(","):
            for lag in a.lags.split(","):
                for cta in a.ctas.split(","):
                    for ch in a.chunkss.split(","):
                      for skip in (a.skips.split(",") if a.skips else [""]):
                        for lib in a.libs.split(","):
                            env = dict(os.environ)
                            if skip:
                                env["NV12EQ_DEBUG_SKIP"] = skip
                            env.pop("NV12EQ_LIB", None)
                            if lib != "default":
                                env["NV12EQ_LIB"] = os.path.join(ROOT, lib)
                            time.sleep(a.cooldown)
                            r = subprocess.run([sys.executable, __file__, "--child", "--op", a.op, "--size", size, "--lag", lag, "--cta", cta,
                                                "--chunks", ch, "--frames", str(a.frames), "--iters", str(a.iters), "--tiles", str(a.tiles),
                                                "--clip", str(a.clip)], env=env, capture_output=True, text=True)
                            out = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                            print(out[-1] if out else json.dumps({"lib": lib, "size": size, "error": (r.stderr or r.stdout)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
