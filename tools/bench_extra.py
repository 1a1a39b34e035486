#!/usr/bin/env python
"""Secondary measurements (GPU box only), one JSON line each:
  --what color   BASELINE configs[4]: BGR -> YUV, equalizeHist(Y), -> BGR on a device-resident batch of packed BGR frames
  --what stream  BASELINE configs[2]/[3]: CLAHE (or equalizeHist) on a paced NV12 stream through nv12eq_stream_*: sustained
                 frames/s with a producer and a consumer thread, and push->pop latency per frame
Same timing rules as bench.py (CUDA events on the launching stream for the device leg, >= 3 warm-up steps, inputs larger
than L2)."""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import opencv_opencl_b200 as nv12eq  # noqa: E402

SIZES = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def color(args):
    from oracle import oracle as O
    W, H = SIZES[args.size]
    n = args.frames
    pitch = 3 * W * H
    ctx = nv12eq.Context(0, W, H, 1)
    st = torch.cuda.current_stream()
    d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_in)
    ctx.synth_bgr_device(d_in, n, pitch, W, H, first_frame=0, stream=st)
    mode = nv12eq.COLOR_YUV if args.color_mode == "yuv" else nv12eq.COLOR_YCRCB

    def step():
        if args.op == "clahe":
            ctx.color_clahe_device(d_in, d_out, n, pitch, W, H, args.clip, (args.tiles, args.tiles), color_mode=mode, stream=st)
        else:
            ctx.color_equalize_device(d_in, d_out, n, pitch, W, H, color_mode=mode, stream=st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = ctx.counters()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    k = n - 1
    bgr = d_in[k * pitch:(k + 1) * pitch].cpu().numpy().reshape(H, W, 3)
    want = O.c_color_equalize(bgr, mode, use_clahe=(args.op == "clahe"), clip=args.clip, tx=args.tiles, ty=args.tiles)
    ok = bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy().reshape(H, W, 3), want))
    peak, src = peak_gbs()
    gbs = n * 6 * W * H / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "bgr_frames_per_sec", "value": n / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                      "config": {"workload": f"BGR->{args.color_mode.upper()}, {args.op}(Y), ->BGR on a {n}-frame {W}x{H} packed BGR batch",
                                 "frames": n}, "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                                            "frac": gbs / peak, "algorithmic_bytes_per_frame": 6 * W * H,
                                                            "peak_source": src},
                      "gpu_launches": ctx.counters()["kernel_launches"] - l0, "parity_spot_check": ok, "steps": args.steps}))


def stream(args):
    from oracle import oracle as O
    W, H = SIZES[args.size]
    ctxs = [nv12eq.Context(g, W, H, 1) for g in range(args.gpus)]   # one context per GPU, one process (config 4 shape)
    op = nv12eq.OP_CLAHE if args.op == "clahe" else nv12eq.OP_EQUALIZE
    frames = [O.c_synth_nv12(W, H, 2026, k) for k in range(8)]
    want = [O.c_nv12_clahe(f, W, H, args.clip, args.tiles, args.tiles) if args.op == "clahe" else O.c_nv12_equalize_hist(f, W, H)
            for f in frames]
    n = args.frames
    res = {}
    for label, fps in (("paced", args.fps), ("unpaced", 0)):
        subs = [nv12eq.Stream(c, W, H, op=op, clip_limit=args.clip, tiles=(args.tiles, args.tiles), depth=args.depth,
                              full_policy=nv12eq.FULL_BLOCK) for c in ctxs]
        s = subs[0] if len(subs) == 1 else nv12eq.sharding.FrameShardedStream(subs)   # frame k -> GPU k mod N, in-order pop
        lat, bad, out = [], [0], np.empty(subs[0].frame_bytes, np.uint8)
        t_push = {}

        def producer():
            t0 = time.perf_counter()
            for k in range(n):
                if fps:
                    due = t0 + k / fps
                    while time.perf_counter() < due:
                        time.sleep(0.0002)
                t_push[k] = time.perf_counter()
                s.push(frames[k % 8])

        def consumer():
            for k in range(n):
                q, f = s.pop(out=out, block=True, wait_push=True) if len(subs) > 1 else s.pop(out=out, block=True)
                lat.append(time.perf_counter() - t_push[q])
                if q != k or (k % 16 == 0 and not np.array_equal(f, want[q % 8])):
                    bad[0] += 1
        tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
        t0 = time.perf_counter()
        tp.start(); tc.start(); tp.join(); tc.join()
        dt = time.perf_counter() - t0
        sts = [x.stats() for x in subs]
        st = {"max_in_flight": sum(x["max_in_flight"] for x in sts), "dropped_backpressure": sum(x["dropped_backpressure"] for x in sts)}
        s.close()
        lat_ms = np.array(lat) * 1e3
        res[label] = {"target_fps": fps or None, "frames_per_sec": n / dt, "latency_ms_p50": float(np.percentile(lat_ms, 50)),
                      "latency_ms_p99": float(np.percentile(lat_ms, 99)), "latency_ms_max": float(lat_ms.max()),
                      "in_order_and_bit_exact": bad[0] == 0, "max_in_flight": st["max_in_flight"], "dropped": st["dropped_backpressure"]}
    print(json.dumps({"metric": "nv12_stream", "config": {"workload": f"{args.op} on a {W}x{H} NV12 stream through nv12eq_stream_* "
                                                                      f"(host frames in, host frames out, depth {args.depth} per GPU, "
                                                                      f"{args.gpus} GPU(s) in one process, frame k -> GPU k mod N)",
                                                          "frames": n}, **res}))


def clahe16(args):
    """16-bit CLAHE (CV_16UC1 / P010 luma, 65536-bin path) on device-resident planes."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from cases_ext import plane16
    from oracle import oracle as O
    W, H = SIZES[args.size]
    n = min(args.frames, 32)
    ctx = nv12eq.Context(0, W, H, 1)
    st = torch.cuda.current_stream()
    base = np.stack([plane16(W, H, args.kind, 50 + k) for k in range(4)])
    d_in = torch.from_numpy(np.concatenate([base] * (n // 4)).view(np.int16)).cuda()
    d_out = torch.zeros_like(d_in)

    def step():
        ctx.clahe16_device(d_in, d_out, n, W * H, W, H, args.clip, (args.tiles, args.tiles), stream=st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ok = bool(np.array_equal(d_out[1].cpu().numpy().view(np.uint16), O.c_clahe16(base[1], args.clip, args.tiles, args.tiles)))
    peak, src = peak_gbs()
    gbs = n * 4 * W * H / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "u16_planes_per_sec", "value": n / (ms * 1e-3), "unit": "planes/s", "ms_per_step": ms,
                      "config": {"workload": f"CLAHE clip={args.clip} tiles={args.tiles}x{args.tiles} on {n} {W}x{H} CV_16UC1 planes "
                                             + ("(P010-like luma: 10 significant bits)" if args.kind == "p010" else "(uniform random over all 65536 values)")},
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "algorithmic_bytes_per_plane": 4 * W * H, "peak_source": src,
                                   "note": "bound by shared-memory atomics (histogram) and the per-pixel table gather, not by HBM"},
                      "parity_spot_check": ok, "steps": args.steps}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", required=True, choices=["color", "stream", "clahe16"])
    ap.add_argument("--op", default="equalize", choices=["equalize", "clahe"])
    ap.add_argument("--size", default="4k", choices=sorted(SIZES))
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--clip", type=float, default=2.0)
    ap.add_argument("--tiles", type=int, default=8)
    ap.add_argument("--color-mode", default="yuv", choices=["yuv", "ycrcb"])
    ap.add_argument("--fps", type=float, default=60.0)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--kind", default="p010", choices=["p010", "full"], help="clahe16: content of the synthetic planes")
    ap.add_argument("--gpus", type=int, default=1, help="stream mode: GPUs driven by this one process")
    args = ap.parse_args()
    nv12eq.build()
    {"color": color, "stream": stream, "clahe16": clahe16}[args.what](args)


if __name__ == "__main__":
    main()
