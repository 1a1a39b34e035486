#!/usr/bin/env python
"""Generate the golden vectors that pin the oracle (and, through it, the CUDA path) to the reference.

The reference's arithmetic on this path is OpenCV's (`cv::equalizeHist`, `cv::CLAHE::apply`,
`cv::cvtColor`: nextimprovement.cpp:168, clahevideo.cpp:184-195, singlecolor.cpp:39-66); the reference holds
no golden vectors of its own.  This script runs those very functions through OpenCV's Python
bindings (cv2 4.13.0 in the build container) and writes

  tests/golden/golden.json   sha1 digests of cv2 outputs on deterministic inputs (Appendix B generator and
                             seeded NumPy distributions), plus the CLAHE geometry each case resolves to;
  tests/golden/fixtures.npz  small raw input/output arrays (so a digest mismatch can be localised).

Run from the repo root:  python tests/golden/make_golden.py
Only needed when cases are added; the outputs are committed.  `hun.png` (the reference's one real image) is
hashed when /root/reference is mounted; the image itself is not copied into this repo.
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


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


from cases import (CLAHE_PARAMS, DIST_KINDS, DIST_SIZES, SYNTH_SIZES, dist_image)  # noqa: E402


def clahe_or_none(y, clip, tx, ty):
    try:
        return cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(y)
    except cv2.error:
        return None


def main():
    golden = {"cv2_version": cv2.__version__, "synth": [], "dist": [], "color": [], "hun": None}
    fixtures = {}

    for (W, H) in SYNTH_SIZES:
        for frame in (0, 1):
            nv = O.c_synth_nv12(W, H, 2026, frame)
            assert (nv == O.np_synth_nv12(W, H, 2026, frame)).all()
            y = nv[:W * H].reshape(H, W)
            rec = {"W": W, "H": H, "seed": 2026, "frame": frame, "in_y": sha(y), "in_uv": sha(nv[W * H:]),
                   "eq": sha(cv2.equalizeHist(y)), "clahe": {}}
            params = CLAHE_PARAMS if W * H <= 1920 * 1080 else CLAHE_PARAMS[:3]
            for (clip, tx, ty) in params:
                out = clahe_or_none(y, clip, tx, ty)
                if out is not None:
                    rec["clahe"][f"{clip}:{tx}:{ty}"] = sha(out)
            golden["synth"].append(rec)
            if (W, H) in ((64, 48), (34, 18), (16, 2)) and frame == 0:
                fixtures[f"synth_{W}x{H}_in"] = y.copy()
                fixtures[f"synth_{W}x{H}_eq"] = cv2.equalizeHist(y)
                fixtures[f"synth_{W}x{H}_clahe_2.0_8_8"] = clahe_or_none(y, 2.0, 8, 8)

    seed = 1000
    for kind in DIST_KINDS:
        for (W, H) in DIST_SIZES:
            seed += 1
            y = dist_image(kind, W, H, seed)
            rec = {"kind": kind, "W": W, "H": H, "seed": seed, "in": sha(y), "eq": sha(cv2.equalizeHist(y)),
                   "clahe": {}}
            for (clip, tx, ty) in CLAHE_PARAMS:
                out = clahe_or_none(y, clip, tx, ty)
                if out is not None:
                    rec["clahe"][f"{clip}:{tx}:{ty}"] = sha(out)
            golden["dist"].append(rec)
            if (W, H) in ((61, 47), (7, 3)):
                fixtures[f"dist_{kind}_{W}x{H}_in"] = y
                fixtures[f"dist_{kind}_{W}x{H}_eq"] = cv2.equalizeHist(y)
                out = clahe_or_none(y, 2.0, 7, 5)
                if out is not None:
                    fixtures[f"dist_{kind}_{W}x{H}_clahe_2.0_7_5"] = out

    for (W, H) in [(3840, 2160), (1920, 1080), (322, 200), (31, 9)]:
        bgr = O.c_synth_bgr(W, H, 0)
        rec = {"W": W, "H": H, "in": sha(bgr)}
        for name, mode in (("yuv", O.COLOR_YUV), ("ycrcb", O.COLOR_YCRCB)):
            fwd = cv2.COLOR_BGR2YUV if mode == O.COLOR_YUV else cv2.COLOR_BGR2YCrCb
            ycc = cv2.cvtColor(bgr, fwd)
            rec[f"{name}_fwd"] = sha(ycc)
            rec[f"{name}_eq"] = sha(O.cv2_color_equalize(bgr, mode))
            rec[f"{name}_clahe_3.0_4_4"] = sha(O.cv2_color_equalize(bgr, mode, True, 3.0, 4, 4))
        golden["color"].append(rec)
        if (W, H) == (31, 9):
            fixtures["color_31x9_in"] = bgr
            fixtures["color_31x9_yuv_eq"] = O.cv2_color_equalize(bgr, O.COLOR_YUV)
            fixtures["color_31x9_ycrcb_eq"] = O.cv2_color_equalize(bgr, O.COLOR_YCRCB)

    # Full 2^24 BGR cube through both conversions (digest only).
    cube = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([(cube & 255), (cube >> 8) & 255, (cube >> 16) & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    golden["cube"] = {
        "bgr2yuv": sha(cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV)), "yuv2bgr": sha(cv2.cvtColor(bgr, cv2.COLOR_YUV2BGR)),
        "bgr2ycrcb": sha(cv2.cvtColor(bgr, cv2.COLOR_BGR2YCrCb)),
        "ycrcb2bgr": sha(cv2.cvtColor(bgr, cv2.COLOR_YCrCb2BGR))}

    hun = "/root/reference/hun.png"
    if os.path.exists(hun):
        img = cv2.imread(hun)
        y = cv2.cvtColor(img, cv2.COLOR_BGR2YUV)[..., 0].copy()
        golden["hun"] = {"shape": list(y.shape), "y": sha(y), "eq": sha(cv2.equalizeHist(y)),
                         "clahe_2.0_8_8": sha(cv2.createCLAHE(2.0, (8, 8)).apply(y)),
                         "color_yuv_eq": sha(O.cv2_color_equalize(img, O.COLOR_YUV))}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **fixtures)
    print("wrote", len(golden["synth"]), "synth,", len(golden["dist"]), "dist,", len(golden["color"]), "color cases;",
          len(fixtures), "fixture arrays")


if __name__ == "__main__":
    main()
