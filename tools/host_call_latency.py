import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
import opencv_opencl_b200 as nv
from oracle import oracle as O
for (W, H) in ((3840, 2160), (1920, 1080)):
    fr = O.c_synth_nv12(W, H, 2026, 0)
    out = np.empty_like(fr)
    with nv.Context(0, W, H, 2) as ctx:
        for op in ("equalize", "clahe"):
            f = (lambda: ctx.equalize_hist(fr, W, H, out=out)) if op == "equalize" else (lambda: ctx.clahe(fr, W, H, 2.0, (8, 8), out=out))
            for _ in range(5): f()
            t = []
            for _ in range(40):
                t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
            t.sort()
            want = O.c_nv12_equalize_hist(fr, W, H) if op == "equalize" else O.c_nv12_clahe(fr, W, H, 2.0, 8, 8)
            print(f"threads={os.environ.get('NV12EQ_HOST_THREADS','default')} {W}x{H} {op}: pageable frame in / frame out p50 {t[20]*1e3:.2f} ms p90 {t[36]*1e3:.2f} ms, bit-exact {np.array_equal(out, want)}")
