#!/usr/bin/env python
"""Time the device-resident kernels over a grid of tuning knobs (GPU box only).  Prints one line per configuration:
op size frames chunks lag ctas schedule -> ms per batch, us per frame, fraction of the measured HBM roofline."""
import argparse
import time
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import opencv_opencl_b200 as nv12eq  # noqa: E402

SIZES = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="equalize,clahe")
    ap.add_argument("--sizes", default="4k")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--lags", default="0")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--schedules", default="0")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tiles", type=int, default=8, help="CLAHE tile grid (tiles x tiles)")
    ap.add_argument("--cooldown", type=float, default=1.0, help="idle seconds before each configuration")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    nv12eq.build()
    st = torch.cuda.current_stream()
    for size in args.sizes.split(","):
        W, H = SIZES[size]
        n = args.frames
        pitch = nv12eq.nv12_frame_bytes(W, H)
        ctx = nv12eq.Context(0, W, H, 1)
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
        d_out = torch.empty_like(d_in)
        ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
        for op in args.ops.split(","):
            for c, l, k, s in itertools.product(*[[int(x) for x in v.split(",")] for v in (args.chunks, args.lags, args.ctas, args.schedules)]):
                ctx.set_tuning(c, l, k, s)
                time.sleep(args.cooldown)   # back-to-back configurations push the board into its power cap: later ones would look slower

                def step():
                    if op == "equalize":
                        ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=st)
                    else:
                        ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (args.tiles, args.tiles), stream=st)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(args.iters):
                    step()
                e1.record(st)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                gbs = n * 3 * W * H / (ms * 1e-3) / 1e9
                print(f"{op:8s} {size:5s} n={n:4d} chunks={c:4d} lag={l:2d} ctas={k} sched={s} : {ms:8.3f} ms/batch "
                      f"{ms * 1e3 / n:7.2f} us/frame {n / (ms * 1e-3):10.0f} fps {gbs:7.0f} GB/s frac={gbs / peak:.3f}", flush=True)
        ctx.close()
        del d_in, d_out


if __name__ == "__main__":
    main()
