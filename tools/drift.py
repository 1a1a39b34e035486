#!/usr/bin/env python
"""Per-launch time of one workload over many back-to-back launches, with NVML clocks / power sampled alongside: shows
whether throughput drifts under sustained load (GPU box only)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import opencv_opencl_b200 as nv
op = sys.argv[1] if len(sys.argv) > 1 else "clahe"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
W, H, n = 3840, 2160, 256
pitch = nv.nv12_frame_bytes(W, H)
c = nv.Context(0, W, H, 1); st = torch.cuda.current_stream()
a = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
c.synth_nv12_device(a, n, pitch, W, H, stream=st)
f = {"equalize": lambda: c.equalize_hist_device(a, b, n, pitch, W, H, stream=st),
     "clahe": lambda: c.clahe_device(a, b, n, pitch, W, H, 2.0, (8, 8), stream=st),
     "copy": lambda: b.copy_(a)}[op]   # plain device copy of the same bytes: what does bandwidth alone cost in power?
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], threading.Event()
def sampler():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetTemperature(h, 0), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        stop.wait(0.01)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
t = threading.Thread(target=sampler); t.start()
t0 = time.perf_counter()
ev[0].record(st)
for i in range(iters):
    f(); ev[i + 1].record(st)
torch.cuda.synchronize(); t1 = time.perf_counter()
stop.set(); t.join()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
for i in range(0, iters, max(1, iters // 20)):
    blk = ms[i:i + max(1, iters // 20)]
    print(f"launch {i:4d}: {sum(blk) / len(blk):.3f} ms")
print("elapsed", round(t1 - t0, 3), "s; samples (t, sm MHz, mem MHz, W, C, reasons):")
for s in samples[:: max(1, len(samples) // 12)]:
    print(f"  {s[0] - t0:6.3f}s sm {s[1]} mem {s[2]} {s[3]:.0f} W {s[4]} C reasons {s[5]:#x}")
