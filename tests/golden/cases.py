"""Deterministic case definitions shared by make_golden.py (cv2 side) and the tests (oracle / CUDA side)."""
import numpy as np


def dist_image(kind, W, H, seed):
    """Seeded small-image distributions (PCG64 streams are stable across NumPy versions)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == "normal":
        return np.clip(rng.normal(120, 25, (H, W)), 0, 255).astype(np.uint8)
    if kind == "constant":
        return np.full((H, W), int(rng.integers(0, 256)), np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    if kind == "lowcontrast":
        return (100 + rng.integers(0, 9, (H, W))).astype(np.uint8)
    if kind == "ramp":
        return ((np.arange(W)[None, :] * 3 + np.arange(H)[:, None] * 5) % 256).astype(np.uint8)
    if kind == "onehot":  # one pixel differs from a constant background
        a = np.full((H, W), 17, np.uint8)
        a[H // 2, W // 2] = 203
        return a
    raise ValueError(kind)


SYNTH_SIZES = [(1920, 1080), (3840, 2160), (1280, 720), (1918, 1078), (640, 480), (64, 48), (34, 18), (16, 2),
               (1000, 600), (4096, 2304)]
CLAHE_PARAMS = [(2.0, 8, 8), (3.0, 4, 4), (40.0, 8, 8), (0.01, 8, 8), (0.0, 8, 8), (2.0, 7, 5), (2.0, 1, 1),
                (4.0, 16, 16), (2.0, 3, 2)]
DIST_KINDS = ["uniform", "normal", "constant", "binary", "lowcontrast", "ramp", "onehot"]
DIST_SIZES = [(1, 1), (2, 1), (7, 3), (16, 16), (61, 47), (128, 96), (250, 130), (333, 77)]
