#!/usr/bin/env python
"""Golden vectors for the NV12 <-> BGR adapters.  Outputs come from cv2 (4.13.0 in the build
container): COLOR_YUV2BGR_NV12 directly; BGR -> NV12 as COLOR_BGR2YUV_I420 with the two chroma planes re-interleaved (OpenCV
has no direct BGR -> NV12 code).  Inputs: the Appendix B generators and seeded NumPy streams.

Run from the repo root:  python tests/golden/make_golden_nv12bgr.py   ->  tests/golden/golden_nv12bgr.json, fixtures_nv12bgr.npz
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (input generator only; outputs below come from cv2)

HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_SIZES = [(3840, 2160), (1920, 1080), (1280, 720), (322, 200), (64, 48), (6, 4), (2, 2)]
RANDOM = [(250, 130, 21), (18, 10, 22), (4, 2, 23), (1918, 1078, 24)]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def i420_to_nv12(i420, W, H):
    flat = i420.reshape(-1)
    u, v = flat[W * H:W * H + W * H // 4], flat[W * H + W * H // 4:]
    return np.concatenate([flat[:W * H], np.stack([u, v], -1).reshape(-1)])


def main():
    golden = {"cv2_version": cv2.__version__, "synth": [], "random": []}
    fixtures = {}
    for (W, H) in SYNTH_SIZES:
        nv12 = O.c_synth_nv12(W, H, 2026, 0)
        # the synthetic frame has neutral chroma: add structure so that every term of the conversion is exercised
        nv12 = nv12.copy()
        nv12[W * H:] = O.c_synth_nv12(W, H, 7026, 1)[:W * H // 2]
        bgr = cv2.cvtColor(nv12.reshape(H * 3 // 2, W), cv2.COLOR_YUV2BGR_NV12)
        src = O.c_synth_bgr(W, H, 0)
        back = i420_to_nv12(cv2.cvtColor(src, cv2.COLOR_BGR2YUV_I420), W, H)
        golden["synth"].append({"W": W, "H": H, "nv12_in": sha(nv12), "bgr": sha(bgr), "bgr_in": sha(src), "nv12": sha(back)})
        if (W, H) == (6, 4):
            fixtures["nv12_6x4_in"], fixtures["bgr_6x4_out"] = nv12, bgr
            fixtures["bgr_6x4_in"], fixtures["nv12_6x4_out"] = src, back
    for (W, H, seed) in RANDOM:
        rng = np.random.default_rng(seed)
        nv12 = rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8)
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        golden["random"].append({"W": W, "H": H, "seed": seed, "nv12_in": sha(nv12),
                                 "bgr": sha(cv2.cvtColor(nv12.reshape(H * 3 // 2, W), cv2.COLOR_YUV2BGR_NV12)), "bgr_in": sha(src),
                                 "nv12": sha(i420_to_nv12(cv2.cvtColor(src, cv2.COLOR_BGR2YUV_I420), W, H))})
    # every (Y, U, V) extreme: the saturation branches of the inverse conversion
    ext = np.array([[y, u, v] for y in (0, 16, 17, 235, 255) for u in (0, 128, 255) for v in (0, 128, 255)], dtype=np.uint8)
    W, H = 2 * len(ext), 2
    nv12 = np.concatenate([np.repeat(ext[:, 0], 2), np.repeat(ext[:, 0], 2), ext[:, 1:].reshape(-1)])
    golden["extremes"] = {"W": W, "H": H, "nv12_in": sha(nv12), "bgr": sha(cv2.cvtColor(nv12.reshape(H * 3 // 2, W), cv2.COLOR_YUV2BGR_NV12))}
    fixtures["extremes_nv12"] = nv12
    with open(os.path.join(HERE, "golden_nv12bgr.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fixtures_nv12bgr.npz"), **fixtures)
    print("wrote", len(golden["synth"]) + len(golden["random"]) + 1, "cases")


if __name__ == "__main__":
    main()
