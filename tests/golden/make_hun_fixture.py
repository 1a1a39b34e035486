#!/usr/bin/env python
"""Commit the reference's one real image as a test fixture, with the answers OpenCV gives on it.

`hun.png` (1919x1079 BGR: odd size, real content, exercises CLAHE's padding path) is the only image the reference holds
(BASELINE.md's parity bar names it).  /root/reference does not exist on the GPU box, so the decoded
pixels are stored here (test infrastructure, not product source) together with digests of what the reference's own
OpenCV calls produce on them (cv2 4.13.0 in the build container):

  tests/golden/hun_bgr.npz     the decoded BGR pixels (np.savez_compressed)
  tests/golden/golden_hun.json sha1 of: the Y plane (cvtColor BGR2YUV, singlecolor.cpp:39), equalizeHist(Y)
                               (singlecolor.cpp:55), CLAHE 2.0/8x8 (clahevideo.cpp:184-195) and 3.0/4x4
                               (clahe1frame.cpp:55-56,88-93) on Y, and the whole colour pipelines
                               BGR->YUV->eq|CLAHE->BGR (singlecolor.cpp:39-66, clahe1frame.cpp:83-102), YUV and YCrCb.

Run from the repo root in the build container:  python tests/golden/make_hun_fixture.py
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (cv2 wrappers only)

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    img = cv2.imread("/root/reference/hun.png")
    assert img is not None and img.shape == (1079, 1919, 3)
    y = cv2.cvtColor(img, cv2.COLOR_BGR2YUV)[..., 0].copy()
    g = {"cv2_version": cv2.__version__, "shape": list(img.shape), "bgr": sha(img), "y": sha(y),
         "eq": sha(cv2.equalizeHist(y)),
         "clahe_2.0_8_8": sha(cv2.createCLAHE(2.0, (8, 8)).apply(y)),
         "clahe_3.0_4_4": sha(cv2.createCLAHE(3.0, (4, 4)).apply(y)),
         "clahe_40.0_8_8": sha(cv2.createCLAHE(40.0, (8, 8)).apply(y)),
         "color_yuv_eq": sha(O.cv2_color_equalize(img, O.COLOR_YUV)),
         "color_ycrcb_eq": sha(O.cv2_color_equalize(img, O.COLOR_YCRCB)),
         "color_yuv_clahe_3.0_4_4": sha(O.cv2_color_equalize(img, O.COLOR_YUV, True, 3.0, 4, 4)),
         "color_yuv_clahe_2.0_8_8": sha(O.cv2_color_equalize(img, O.COLOR_YUV, True, 2.0, 8, 8))}
    old = json.load(open(os.path.join(HERE, "golden.json")))["hun"]
    for k in ("y", "eq", "clahe_2.0_8_8", "color_yuv_eq"):
        assert old[k] == g[k], k     # the digests SURVEY.md Appendix B quotes
    np.savez_compressed(os.path.join(HERE, "hun_bgr.npz"), bgr=img)
    with open(os.path.join(HERE, "golden_hun.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote hun_bgr.npz and golden_hun.json:", g["eq"], g["clahe_2.0_8_8"])


if __name__ == "__main__":
    main()
