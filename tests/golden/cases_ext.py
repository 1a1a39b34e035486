"""Deterministic 16-bit case definitions shared by make_golden_ext.py (cv2 side) and the tests (oracle / CUDA side)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402  (Appendix B generator only)

CLAHE16_CASES = [(1920, 1080, "p010", 21), (640, 360, "p010", 22), (322, 200, "full", 23), (101, 67, "p010", 24),
                 (34, 18, "full", 25), (256, 144, "lowcontrast", 26), (64, 48, "constant", 27), (3840, 2160, "p010", 28)]
CLAHE16_PARAMS = [(2.0, 8, 8), (40.0, 4, 4), (0.0, 8, 8), (2.0, 4, 3), (3.0, 1, 1)]


def plane16(W, H, kind, seed):
    """Deterministic 16-bit planes.  "p010": the Appendix B luma (8 bits) spread to 10 bits with seeded noise, in the high
    bits of the word like a P010 decoder delivers it."""
    rng = np.random.default_rng(seed)
    if kind == "p010":
        y8 = O.c_synth_nv12(W, H, 2026, seed)[:W * H].reshape(H, W).astype(np.uint16)
        return (((y8 << 2) | rng.integers(0, 4, (H, W)).astype(np.uint16)) << 6).astype(np.uint16)
    if kind == "full":
        return rng.integers(0, 65536, (H, W), dtype=np.uint16)
    if kind == "lowcontrast":
        return (30000 + rng.integers(0, 300, (H, W))).astype(np.uint16)
    if kind == "constant":
        return np.full((H, W), 12345, np.uint16)
    raise ValueError(kind)


