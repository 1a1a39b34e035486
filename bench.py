#!/usr/bin/env python
"""bench.py -- NV12 frames/s of the hot path on N B200s (one process per GPU), with roofline, end-to-end and CPU legs.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" is one pass of the hot path over one batch of synthetic NV12 frames (SURVEY.md Appendix B generator):

  default workload  = BASELINE.json configs[1]: cv::equalizeHist semantics on a 256-frame 3840x2160 NV12 batch,
                      device resident (2 x 3.19 GB, far larger than the 126 MB L2, so no L2 flush is needed).
  value             = frames/s, inputs already in HBM, timed with CUDA events on the stream the kernel runs on.
  e2e               = frames/s through the host C-ABI call (nv12eq_equalize_hist_batch) with pinned HOST buffers:
                      host->device and device->host copies are inside the timed region.
  roofline          = algorithmic bytes (3*W*H per frame) / kernel time vs the measured HBM peak.
  cpu_baseline      = the reference's own CPU implementation (OpenCV through cv2, else the C oracle port) on a bounded
                      sample of the same frames, all host cores, rank 0, N=1 only.

`--impl reference` times only that CPU implementation (the reference has no GPU path to run).
Multi-GPU: frames are independent units, sharded over ranks with no collective on the data path (weak scaling: every
rank processes its own 256-frame batch); torch.distributed is used only for the barrier and the max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZES = {"4k": (3840, 2160), "1080p": (1920, 1080), "720p": (1280, 720)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--op", default="equalize", choices=["equalize", "clahe"])
    ap.add_argument("--size", default="4k", choices=sorted(SIZES))
    ap.add_argument("--frames", type=int, default=256, help="frames per batch per GPU")
    ap.add_argument("--clip", type=float, default=2.0)
    ap.add_argument("--tiles", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline leg")
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--lag", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--schedule", type=int, default=0)
    ap.add_argument("--slots", type=int, default=2, help="host lanes of the context (e2e leg pipelining depth)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML (same counters nvidia-smi prints)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (checker code: oracle/ is only ever used here as the reported baseline, never on the product path)
# ------------------------------------------------------------------------------------------------------------
def cpu_leg(args, W, H, n_frames, steps, warmup):
    """Frames/s of the reference's CPU implementation on `n_frames` synthetic frames per step, all host cores.
    cv2 available -> kind 'reference' (the very OpenCV functions the reference calls, frame-parallel like its
    --workers threads, nextimprovement.cpp:159-168 / clahevideo.cpp:178-201); otherwise the C oracle port."""
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    frames = np.stack([O.c_synth_nv12(W, H, 2026, k) for k in range(n_frames)])
    out = np.empty_like(frames)
    if O.have_cv2():
        import cv2
        from concurrent.futures import ThreadPoolExecutor
        cv2.setNumThreads(1)  # one frame per worker thread; cv2 releases the GIL
        workers = min(cores, n_frames)
        local = threading.local()

        def one(k):
            if args.op == "equalize":
                O.cv2_nv12_equalize_hist(frames[k], W, H, out[k])
            else:
                if not hasattr(local, "clahe"):
                    local.clahe = cv2.createCLAHE(clipLimit=args.clip, tileGridSize=(args.tiles, args.tiles))
                O.cv2_nv12_clahe(frames[k], W, H, out[k], clahe=local.clahe)

        pool = ThreadPoolExecutor(max_workers=workers)

        def step():
            list(pool.map(one, range(n_frames)))
        kind, used = "reference", workers
        impl = f"cv2 {cv2.__version__} {'equalizeHist' if args.op == 'equalize' else 'CLAHE.apply'} + UV memcpy, {workers} frame-parallel threads"
    else:
        used = O.max_threads()

        def step():
            O.c_nv12_batch(args.op, frames, W, H, clip=args.clip, tx=args.tiles, ty=args.tiles, threads=0, out=out)
        kind = "port"
        impl = f"C oracle port (oracle/nv12eq_oracle.c), {used} OpenMP threads"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # self-check of the baseline itself against the C oracle on one frame
    want = (O.c_nv12_equalize_hist(frames[0], W, H) if args.op == "equalize"
            else O.c_nv12_clahe(frames[0], W, H, args.clip, args.tiles, args.tiles))
    assert np.array_equal(out[0], want), "CPU baseline disagrees with the oracle"
    return {"value": n_frames * steps / dt, "unit": "frames/s", "cores": used, "host_cores": cores, "kind": kind,
            "sample": f"{n_frames} synthetic {W}x{H} NV12 frames x {steps} passes ({impl})", "seconds": dt,
            "ms_per_step": dt / steps * 1e3}


def calibrate_cpu_frames(args, W, H, target_s, passes):
    """Pick a frame count so that `passes` passes take about target_s seconds of wall time."""
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    nv12 = O.c_synth_nv12(W, H, 2026, 0)
    out = np.empty_like(nv12)
    def one():
        if O.have_cv2():
            import cv2
            cv2.setNumThreads(1)
            if args.op == "equalize":
                O.cv2_nv12_equalize_hist(nv12, W, H, out)
            else:
                O.cv2_nv12_clahe(nv12, W, H, out, clip=args.clip, tx=args.tiles, ty=args.tiles)
        else:
            O.c_nv12_batch(args.op, nv12[None], W, H, clip=args.clip, tx=args.tiles, ty=args.tiles, threads=1)
    one()                      # first call: imports, page faults, thread pools
    t0 = time.perf_counter()
    one()
    per_frame = max(time.perf_counter() - t0, 1e-4)
    n = int(target_s / passes / per_frame * min(cores, 16))
    n = max(min(cores, 128), min(n, 128))
    calibrate_cpu_frames.per_frame_s = per_frame   # single-thread seconds per frame, for sizing the number of passes
    return max(4, n)


def workload_name(args, W, H):
    op = "equalizeHist" if args.op == "equalize" else f"CLAHE clip={args.clip} tiles={args.tiles}x{args.tiles}"
    return f"{op} on a {args.frames}-frame {W}x{H} NV12 batch per GPU (BASELINE configs[1] shape)"


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    W, H = SIZES[args.size]
    passes = args.steps + args.warmup
    n = calibrate_cpu_frames(args, W, H, 60.0, passes)
    r = cpu_leg(args, W, H, n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "nv12_frames_per_sec", "value": r["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, W, H), "op": args.op, "width": W, "height": H,
                   "frames_per_step": n, "note": "CPU arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import opencv_opencl_b200 as nv12eq

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nv12eq has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nv12eq.build()

    W, H = SIZES[args.size]
    n = args.frames
    pitch = nv12eq.nv12_frame_bytes(W, H)
    bytes_per_frame_algo = 3 * W * H  # read NV12 once + write NV12 once
    ctx = nv12eq.Context(device=local, max_width=W, max_height=H, slots=args.slots)
    ctx.set_tuning(args.chunks, args.lag, args.ctas, args.schedule)
    stream = torch.cuda.current_stream()
    d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_in)
    ctx.synth_nv12_device(d_in, n, pitch, W, H, seed=2026, first_frame=rank * n, stream=stream)

    def device_step():
        if args.op == "equalize":
            ctx.equalize_hist_device(d_in, d_out, n, pitch, W, H, stream=stream)
        else:
            ctx.clahe_device(d_in, d_out, n, pitch, W, H, args.clip, (args.tiles, args.tiles), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    launches0 = ctx.counters()["kernel_launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local) as clocks:
        ev[0].record(stream)
        for i in range(args.steps):
            device_step()
            ev[i + 1].record(stream)
        barrier()
    launches = ctx.counters()["kernel_launches"] - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    value = world * n * args.steps / (total_ms * 1e-3)
    kernel_ms = sum(step_ms) / len(step_ms)  # one kernel launch per step: its duration is the step's
    achieved = n * bytes_per_frame_algo / (kernel_ms * 1e-3) / 1e9

    # spot check of the timed output against the oracle (rank 0, one frame) -- a fast wrong kernel is not done
    parity = None
    if rank == 0:
        from oracle import oracle as O
        k = n - 1
        frame = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
        want = (O.c_nv12_equalize_hist(frame, W, H) if args.op == "equalize"
                else O.c_nv12_clahe(frame, W, H, args.clip, args.tiles, args.tiles))
        parity = bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), want))

    # ---- end-to-end leg: host C-ABI call with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e_n = n
        try:
            h_in, h_out = nv12eq.PinnedBuffer(e2e_n * pitch), nv12eq.PinnedBuffer(e2e_n * pitch)
        except nv12eq.Nv12eqError:
            e2e_n = max(8, n // 8)
            h_in, h_out = nv12eq.PinnedBuffer(e2e_n * pitch), nv12eq.PinnedBuffer(e2e_n * pitch)
        torch.cuda.synchronize()
        # fill the pinned input from the device copy of the same synthetic frames (outside the timed region)
        import ctypes
        torch.from_numpy(h_in.array).copy_(d_in[:e2e_n * pitch])
        torch.cuda.synchronize()

        def host_step():
            if args.op == "equalize":
                ctx.equalize_hist_batch(h_in.array, W, H, out=h_out.array, n_frames=e2e_n, frame_pitch=pitch)
            else:
                ctx.clahe_batch(h_in.array, W, H, args.clip, (args.tiles, args.tiles), out=h_out.array, n_frames=e2e_n,
                                frame_pitch=pitch)
        e2e_steps = max(2, min(args.steps, 5))
        host_step()
        barrier()
        c0 = ctx.counters()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        c1 = ctx.counters()
        ok = True
        if rank == 0:
            ok = bool(np.array_equal(h_out.array[(e2e_n - 1) * pitch:e2e_n * pitch],
                                     d_out[(e2e_n - 1) * pitch:e2e_n * pitch].cpu().numpy()))
        # bytes that actually crossed PCIe, from the library's own counters: in passthrough mode only the luma planes
        # move (W*H of the 1.5*W*H bytes of a frame, each way); the chroma is copied host-to-host by the library's
        # host threads, as the reference does with memcpy (nextimprovement.cpp:160)
        e2e = {"value": world * e2e_n * e2e_steps / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": (c1["bytes_in"] - c0["bytes_in"]) // e2e_steps,
               "d2h_bytes_per_step": (c1["bytes_out"] - c0["bytes_out"]) // e2e_steps,
               "host_frame_bytes_per_step": e2e_n * pitch, "frames_per_step": e2e_n, "steps": e2e_steps,
               "api": "nv12eq_equalize_hist_batch" if args.op == "equalize" else "nv12eq_clahe_batch",
               "host_memory": "pinned (nv12eq_host_alloc)", "chroma": "host memcpy inside the call (never crosses PCIe)",
               "matches_device_leg": ok}
        h_in.free(); h_out.free()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback (B200_PROFILING.md): 6650 GB/s"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        traffic = t.get(f"{args.op}_{args.size}_{n}", None)
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu:
        nf = calibrate_cpu_frames(args, W, H, args.cpu_seconds, 3)
        # about --cpu-seconds (default 12) of CPU work in total: the frame count is capped by host memory, so the number of
        # passes over the sample makes up the rest
        passes = int(max(2, min(40, round(args.cpu_seconds / (nf * calibrate_cpu_frames.per_frame_s)))))
        r = cpu_leg(args, W, H, nf, passes, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": "nv12_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, W, H), "op": args.op, "width": W, "height": H, "frames_per_gpu": n,
                   "uv": "passthrough", "l2": f"inputs {n * pitch / 1e9:.2f} GB per GPU > 126 MB L2, no flush needed",
                   "parallelism": f"frame-sharded x{world}, no collective", "tuning": {"chunks": args.chunks, "lag": args.lag,
                                                                                      "ctas": args.ctas, "schedule": args.schedule}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "equalize_kernel" if args.op == "equalize" else "clahe_kernel",
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": n * bytes_per_frame_algo},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        "parity_spot_check": parity, "step_ms_min": min(step_ms), "step_ms_max": max(step_ms),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner, build output)
    # is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    import builtins
    real_print = builtins.print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            real_print(*a, **k)
    builtins.print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in lines:
        real_print(line, flush=True)


if __name__ == "__main__":
    main()
