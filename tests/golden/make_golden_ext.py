#!/usr/bin/env python
"""Golden vectors for the adapters added after the first set (kept separate so golden.json never has to be regenerated).

COLOR_BGR2YUV_I420 is the conversion 1frameMeasure.cpp:32 runs in front of the Y-plane operator; the outputs below come
from cv2 (4.13.0 in the build container), the inputs from the Appendix B generator and seeded NumPy streams.

Run from the repo root:  python tests/golden/make_golden_ext.py   ->  tests/golden/golden_ext.json, fixtures_ext.npz
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
sys.path.insert(0, HERE)
from cases_ext import CLAHE16_CASES, CLAHE16_PARAMS, plane16  # noqa: E402
I420_SYNTH_SIZES = [(3840, 2160), (1920, 1080), (1280, 720), (322, 200), (64, 48), (6, 4), (2, 2)]
I420_RANDOM = [(250, 130, 11), (18, 10, 12), (4, 2, 13)]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    golden = {"cv2_version": cv2.__version__, "i420_synth": [], "i420_random": []}
    fixtures = {}
    for (W, H) in I420_SYNTH_SIZES:
        bgr = O.c_synth_bgr(W, H, 0)
        out = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420)
        # the 1frameMeasure shape: Y plane of the I420 image through equalizeHist
        y = out[:H].copy()
        golden["i420_synth"].append({"W": W, "H": H, "in": sha(bgr), "i420": sha(out), "y_eq": sha(cv2.equalizeHist(y))})
        if (W, H) == (6, 4):
            fixtures["i420_6x4_in"] = bgr
            fixtures["i420_6x4_out"] = out
    for (W, H, seed) in I420_RANDOM:
        bgr = np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)
        golden["i420_random"].append({"W": W, "H": H, "seed": seed, "in": sha(bgr), "i420": sha(cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420))})
    # the eight corners of the BGR cube (range extremes of the Q20 formulas)
    corners = np.array([[[b, g, r] for b in (0, 255) for g in (0, 255)] for r in (0, 255)], dtype=np.uint8).reshape(2, 4, 3)
    golden["i420_corners"] = sha(cv2.cvtColor(corners, cv2.COLOR_BGR2YUV_I420))
    # CLAHE on CV_16UC1 (OpenCV's 65536-bin path): P010-like (10 bits << 6), full-range and low-contrast planes
    golden["clahe16"] = []
    for (W, H, kind, seed) in CLAHE16_CASES:
        y = plane16(W, H, kind, seed)
        rec = {"W": W, "H": H, "kind": kind, "seed": seed, "in": sha(y), "out": {}}
        for (clip, tx, ty) in CLAHE16_PARAMS:
            rec["out"][f"{clip}:{tx}:{ty}"] = sha(cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(y))
        golden["clahe16"].append(rec)
        if (W, H) == (34, 18):
            fixtures["clahe16_34x18_in"] = y
            fixtures["clahe16_34x18_2.0_4_3"] = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(4, 3)).apply(y)
    with open(os.path.join(HERE, "golden_ext.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fixtures_ext.npz"), **fixtures)
    print("wrote", len(golden["i420_synth"]) + len(golden["i420_random"]), "I420 cases")


if __name__ == "__main__":
    main()
