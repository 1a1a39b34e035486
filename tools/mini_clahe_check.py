"""Developer tool: tiny CLAHE / equalizeHist parity check against the oracle with whatever library NV12EQ_LIB selects."""
import sys, numpy as np
sys.path.insert(0, '.')
import opencv_opencl_b200 as nv
from oracle import oracle as O
W, H = 640, 360
fr = np.stack([O.c_synth_nv12(W, H, 2026, k) for k in range(3)])
with nv.Context(0, W, H, 2) as c:
    try:
        out = c.equalize_hist_batch(fr, W, H)
        print('eq ok', all(np.array_equal(out[k], O.c_nv12_equalize_hist(fr[k], W, H)) for k in range(3)))
    except Exception as e:
        print('eq FAILED', str(e)[:150])
with nv.Context(0, W, H, 2) as c:
    try:
        out = c.clahe_batch(fr, W, H, 2.0, (8, 8))
        print('clahe ok', all(np.array_equal(out[k], O.c_nv12_clahe(fr[k], W, H, 2.0, 8, 8)) for k in range(3)))
    except Exception as e:
        print('clahe FAILED', str(e)[:150])
