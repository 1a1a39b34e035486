"""CPU oracle for the NV12 equalizeHist / CLAHE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product (``opencv-opencl_b200``) never does and has no CPU fallback.

Three witnesses live here:

* ``c_*``   -- ctypes bindings of ``oracle/nv12eq_oracle.c`` (the plain-C restatement of the algorithm;
               see that file's header for the reference file:line each function follows).
* ``np_*``  -- a NumPy restatement of SURVEY.md Appendix A, kept as an independent second witness.
* ``cv2_*`` -- the functions the reference itself calls (``cv::equalizeHist``, ``cv::CLAHE::apply``,
               ``cv::cvtColor``; e.g. nextimprovement.cpp:159-168, clahevideo.cpp:184-201,
               singlecolor.cpp:39-66) reached through OpenCV's Python bindings, used (a) by
               ``tests/golden/make_golden.py`` to pin the oracle and (b) as the reference CPU arm of
               ``bench.py`` when ``cv2`` is importable on the box.

Parity status: PINNED -- against committed cv2-generated digests/fixtures in ``tests/golden`` and
against live cv2 wherever it imports.  The reference ships no golden vectors of its own.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnv12eq_oracle.so")

UV_COPY, UV_GRAY128, UV_SKIP = 0, 1, 2
COLOR_YUV, COLOR_YCRCB = 0, 1

_u8p = ctypes.POINTER(ctypes.c_uint8)
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "nv12eq_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_int, c_u32, c_dbl, c_sz = ctypes.c_int, ctypes.c_uint32, ctypes.c_double, ctypes.c_size_t
        L.oracle_synth_nv12.argtypes = [_u8p, c_int, c_int, c_int, c_u32, c_u32]
        L.oracle_synth_nv12.restype = None
        L.oracle_synth_y.argtypes = [_u8p, c_int, c_int, c_int, c_u32, c_u32]
        L.oracle_synth_y.restype = None
        L.oracle_synth_bgr.argtypes = [_u8p, c_int, c_int, c_int, c_u32]
        L.oracle_synth_bgr.restype = None
        L.oracle_hist256.argtypes = [_u8p, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32)]
        L.oracle_hist256.restype = None
        L.oracle_equalize_lut.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int64, _u8p]
        L.oracle_equalize_lut.restype = c_int
        L.oracle_equalize_hist.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int]
        L.oracle_equalize_hist.restype = None
        L.oracle_clahe.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int, c_dbl, c_int, c_int]
        L.oracle_clahe.restype = c_int
        L.oracle_clahe_tile_luts.argtypes = [_u8p, c_int, c_int, c_int, c_dbl, c_int, c_int, _u8p]
        L.oracle_clahe_tile_luts.restype = None
        L.oracle_clahe_geometry.argtypes = [c_int, c_int, c_dbl, c_int, c_int] + [ctypes.POINTER(c_int)] * 5
        L.oracle_clahe_geometry.restype = None
        L.oracle_nv12_equalize_hist.argtypes = [_u8p, _u8p, c_int, c_int, c_int, c_int]
        L.oracle_nv12_equalize_hist.restype = c_int
        L.oracle_nv12_clahe.argtypes = [_u8p, _u8p, c_int, c_int, c_int, c_dbl, c_int, c_int, c_int]
        L.oracle_nv12_clahe.restype = c_int
        L.oracle_nv12_equalize_hist_batch.argtypes = [_u8p, _u8p, c_int, c_sz, c_int, c_int, c_int, c_int, c_int]
        L.oracle_nv12_equalize_hist_batch.restype = c_int
        L.oracle_nv12_clahe_batch.argtypes = [_u8p, _u8p, c_int, c_sz, c_int, c_int, c_int, c_dbl, c_int, c_int,
                                              c_int, c_int]
        L.oracle_nv12_clahe_batch.restype = c_int
        L.oracle_max_threads.argtypes = []
        L.oracle_max_threads.restype = c_int
        L.oracle_bgr2ycc.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int, c_int]
        L.oracle_bgr2ycc.restype = None
        L.oracle_ycc2bgr.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int, c_int]
        L.oracle_ycc2bgr.restype = None
        u16p = ctypes.POINTER(ctypes.c_uint16)
        L.oracle_clahe16.argtypes = [u16p, c_int, u16p, c_int, c_int, c_int, c_dbl, c_int, c_int]
        L.oracle_clahe16.restype = c_int
        L.oracle_bgr2i420.argtypes = [_u8p, c_int, _u8p, c_int, c_int]
        L.oracle_bgr2i420.restype = c_int
        L.oracle_clahe_interp_band.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _u8p, c_int]
        L.oracle_clahe_interp_band.restype = c_int
        L.oracle_nv12_to_bgr.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int]
        L.oracle_nv12_to_bgr.restype = c_int
        L.oracle_bgr_to_nv12.argtypes = [_u8p, c_int, _u8p, c_int, c_int, c_int]
        L.oracle_bgr_to_nv12.restype = c_int
        L.oracle_color_equalize.argtypes = [_u8p, _u8p, c_int, c_int, c_int, c_int, c_int, c_dbl, c_int, c_int]
        L.oracle_color_equalize.restype = c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_u8p)


def sha16(a) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def nv12_bytes(W: int, H: int, stride: int | None = None) -> int:
    stride = W if stride is None else stride
    return stride * (H + H // 2)


# --------------------------------------------------------------------------------------------
# C restatement
# --------------------------------------------------------------------------------------------
def c_synth_nv12(W, H, seed=2026, frame=0, stride=None) -> np.ndarray:
    """Appendix B generator.  Returns the flat NV12 buffer (stride*(H+H//2) bytes)."""
    stride = W if stride is None else stride
    buf = np.zeros(nv12_bytes(W, H, stride), dtype=np.uint8)
    lib().oracle_synth_nv12(_p(buf), stride, W, H, seed, frame)
    return buf


def c_synth_bgr(W, H, frame=0) -> np.ndarray:
    img = np.zeros((H, W, 3), dtype=np.uint8)
    lib().oracle_synth_bgr(_p(img), 3 * W, W, H, frame)
    return img


def c_equalize_hist(y: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y)
    H, W = y.shape
    out = np.empty_like(y)
    lib().oracle_equalize_hist(_p(y), W, _p(out), W, W, H)
    return out


def c_hist256(y: np.ndarray) -> np.ndarray:
    """256-bin int32 histogram of a 2-D uint8 plane (a2.1)."""
    y = np.ascontiguousarray(y)
    H, W = y.shape
    hist = np.zeros(256, dtype=np.int32)
    lib().oracle_hist256(_p(y), W, W, H, hist.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return hist


def c_equalize_lut(hist: np.ndarray, total: int) -> np.ndarray:
    """equalizeHist LUT of a 256-bin histogram describing `total` pixels (a2.2).  For a constant image every entry that
    can be read equals the constant."""
    hist = np.ascontiguousarray(hist, dtype=np.int32)
    lut = np.zeros(256, dtype=np.uint8)
    lib().oracle_equalize_lut(hist.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(total), _p(lut))
    return lut


def c_clahe(y: np.ndarray, clip=2.0, tx=8, ty=8) -> np.ndarray:
    y = np.ascontiguousarray(y)
    H, W = y.shape
    out = np.empty_like(y)
    rc = lib().oracle_clahe(_p(y), W, _p(out), W, W, H, float(clip), tx, ty)
    if rc:
        raise RuntimeError(f"oracle_clahe rc={rc}")
    return out


def c_clahe16(y: np.ndarray, clip=2.0, tx=8, ty=8) -> np.ndarray:
    """CLAHE on a CV_16UC1 plane (OpenCV's 16-bit path: 65536 bins)."""
    y = np.ascontiguousarray(y, dtype=np.uint16)
    H, W = y.shape
    out = np.empty_like(y)
    u16p = ctypes.POINTER(ctypes.c_uint16)
    rc = lib().oracle_clahe16(y.ctypes.data_as(u16p), W, out.ctypes.data_as(u16p), W, W, H, float(clip), tx, ty)
    if rc:
        raise RuntimeError(f"oracle_clahe16 rc={rc}")
    return out


def c_clahe_tile_luts(y: np.ndarray, clip=2.0, tx=8, ty=8) -> np.ndarray:
    y = np.ascontiguousarray(y)
    H, W = y.shape
    luts = np.empty((ty * tx, 256), dtype=np.uint8)
    lib().oracle_clahe_tile_luts(_p(y), W, W, H, float(clip), tx, ty, _p(luts))
    return luts


def c_clahe_interp_band(band: np.ndarray, H: int, tx: int, ty: int, y_first: int, luts_halo: np.ndarray, first_tile_row: int) -> np.ndarray:
    """Interpolation of rows [y_first, y_first + band rows) of a W x H frame from a LUT grid with halo (tile rows
    first_tile_row - 1 ..): the per-rank stage of the spatially split single-frame mode."""
    band = np.ascontiguousarray(band)
    rows, W = band.shape
    luts_halo = np.ascontiguousarray(luts_halo).reshape(-1)
    out = np.empty_like(band)
    rc = lib().oracle_clahe_interp_band(_p(band), W, _p(out), W, W, H, tx, ty, y_first, rows, _p(luts_halo), first_tile_row)
    if rc:
        raise ValueError(f"oracle_clahe_interp_band rc={rc}")
    return out


def c_nv12_equalize_hist(nv12: np.ndarray, W, H, stride=None, uv_mode=UV_COPY, out=None) -> np.ndarray:
    stride = W if stride is None else stride
    out = np.zeros_like(nv12) if out is None else out
    rc = lib().oracle_nv12_equalize_hist(_p(nv12), _p(out), W, H, stride, uv_mode)
    if rc:
        raise RuntimeError(f"oracle_nv12_equalize_hist rc={rc}")
    return out


def c_nv12_clahe(nv12: np.ndarray, W, H, clip=2.0, tx=8, ty=8, stride=None, uv_mode=UV_COPY, out=None) -> np.ndarray:
    stride = W if stride is None else stride
    out = np.zeros_like(nv12) if out is None else out
    rc = lib().oracle_nv12_clahe(_p(nv12), _p(out), W, H, stride, float(clip), tx, ty, uv_mode)
    if rc:
        raise RuntimeError(f"oracle_nv12_clahe rc={rc}")
    return out


def c_nv12_batch(op: str, frames: np.ndarray, W, H, clip=2.0, tx=8, ty=8, uv_mode=UV_COPY, threads=0,
                 out=None) -> np.ndarray:
    """frames: (n, nv12_bytes) uint8.  One frame per OpenMP thread (threads=0 -> all cores)."""
    assert frames.ndim == 2
    n, pitch = frames.shape
    out = np.empty_like(frames) if out is None else out
    if op == "equalize":
        rc = lib().oracle_nv12_equalize_hist_batch(_p(frames), _p(out), n, pitch, W, H, W, uv_mode, threads)
    elif op == "clahe":
        rc = lib().oracle_nv12_clahe_batch(_p(frames), _p(out), n, pitch, W, H, W, float(clip), tx, ty, uv_mode,
                                           threads)
    else:
        raise ValueError(op)
    if rc:
        raise RuntimeError(f"oracle batch rc={rc}")
    return out


def c_color_equalize(bgr: np.ndarray, mode=COLOR_YUV, use_clahe=False, clip=2.0, tx=8, ty=8) -> np.ndarray:
    bgr = np.ascontiguousarray(bgr)
    H, W, _ = bgr.shape
    out = np.empty_like(bgr)
    rc = lib().oracle_color_equalize(_p(bgr), _p(out), W, H, 3 * W, mode, int(use_clahe), float(clip), tx, ty)
    if rc:
        raise RuntimeError(f"oracle_color_equalize rc={rc}")
    return out


def c_bgr2i420(bgr: np.ndarray) -> np.ndarray:
    """COLOR_BGR2YUV_I420 (1frameMeasure.cpp:32): returns the (H*3/2, W) planar image exactly as cv2 lays it out."""
    bgr = np.ascontiguousarray(bgr)
    H, W, _ = bgr.shape
    out = np.empty((H * 3 // 2, W), dtype=np.uint8)
    rc = lib().oracle_bgr2i420(_p(bgr), 3 * W, _p(out), W, H)
    if rc:
        raise ValueError("oracle_bgr2i420: width and height must be even")
    return out


def c_nv12_to_bgr(nv12: np.ndarray, W: int, H: int) -> np.ndarray:
    """COLOR_YUV2BGR_NV12 of a flat NV12 frame (W*H*3/2 bytes): returns (H, W, 3) BGR."""
    nv12 = np.ascontiguousarray(nv12).reshape(-1)
    out = np.empty((H, W, 3), dtype=np.uint8)
    rc = lib().oracle_nv12_to_bgr(_p(nv12), W, _p(out), 3 * W, W, H)
    if rc:
        raise ValueError("oracle_nv12_to_bgr: width and height must be even")
    return out


def c_bgr_to_nv12(bgr: np.ndarray) -> np.ndarray:
    """BGR -> flat NV12 frame (COLOR_BGR2YUV_I420 arithmetic, chroma interleaved U first)."""
    bgr = np.ascontiguousarray(bgr)
    H, W, _ = bgr.shape
    out = np.empty(W * H * 3 // 2, dtype=np.uint8)
    rc = lib().oracle_bgr_to_nv12(_p(bgr), 3 * W, _p(out), W, W, H)
    if rc:
        raise ValueError("oracle_bgr_to_nv12: width and height must be even")
    return out


def c_bgr2ycc(bgr: np.ndarray, mode=COLOR_YUV) -> np.ndarray:
    bgr = np.ascontiguousarray(bgr)
    H, W, _ = bgr.shape
    out = np.empty_like(bgr)
    lib().oracle_bgr2ycc(_p(bgr), 3 * W, _p(out), 3 * W, W, H, mode)
    return out


def c_ycc2bgr(ycc: np.ndarray, mode=COLOR_YUV) -> np.ndarray:
    ycc = np.ascontiguousarray(ycc)
    H, W, _ = ycc.shape
    out = np.empty_like(ycc)
    lib().oracle_ycc2bgr(_p(ycc), 3 * W, _p(out), 3 * W, W, H, mode)
    return out


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# --------------------------------------------------------------------------------------------
# NumPy restatement (second witness)
# --------------------------------------------------------------------------------------------
def _fmix32(k: np.ndarray) -> np.ndarray:
    k = k.astype(np.uint32)
    k ^= k >> np.uint32(16)
    k *= np.uint32(0x85EBCA6B)
    k ^= k >> np.uint32(13)
    k *= np.uint32(0xC2B2AE35)
    k ^= k >> np.uint32(16)
    return k


def np_synth_nv12(W, H, seed=2026, frame=0) -> np.ndarray:
    with np.errstate(over="ignore"):
        yy, xx = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
        idx = (yy * W + xx).astype(np.uint32)
        k = _fmix32(idx * np.uint32(0x9E3779B1) + np.uint32((seed * 0x85EBCA77) & 0xFFFFFFFF)
                    + np.uint32((frame * 0xC2B2AE3D) & 0xFFFFFFFF))
        base = 48 + (xx * 128) // W + (yy * 48) // H
        noise = (k & np.uint32(63)).astype(np.int64) - 32
        patch = ((xx // max(W // 16, 1)) + (yy // max(H // 9, 1))) % 5 == 0
        base = np.where(patch, 200, base)
        noise = np.where(patch, (k & np.uint32(3)).astype(np.int64), noise)
        Y = np.clip(base + noise, 0, 255).astype(np.uint8)
        j = np.arange(W * (H // 2), dtype=np.uint32)
        ku = _fmix32(j * np.uint32(0x9E3779B1) + np.uint32((seed + 0x01234567) & 0xFFFFFFFF)
                     + np.uint32((frame * 0xC2B2AE3D) & 0xFFFFFFFF))
        UV = (128 + (ku & np.uint32(31)).astype(np.int64) - 16).astype(np.uint8)
    return np.concatenate([Y.reshape(-1), UV])


def np_equalize_hist(y: np.ndarray) -> np.ndarray:
    hist = np.bincount(y.reshape(-1), minlength=256).astype(np.int64)
    total = y.size
    nz = np.nonzero(hist)[0]
    i0 = int(nz[0])
    if hist[i0] == total:
        return np.full_like(y, i0)
    scale = np.float32(255.0) / np.float32(total - hist[i0])
    cs = np.cumsum(hist) - hist[: i0 + 1].sum()
    lut = np.zeros(256, dtype=np.uint8)
    vals = np.rint(cs[i0 + 1:].astype(np.float32) * scale)  # f32 multiply, round-half-even
    lut[i0 + 1:] = np.clip(vals, 0, 255).astype(np.uint8)
    return lut[y]


def np_clahe(y: np.ndarray, clip=2.0, tx=8, ty=8) -> np.ndarray:
    H, W = y.shape
    if W % tx == 0 and H % ty == 0:
        ext = y
    else:
        ext = np.pad(y, ((0, ty - H % ty), (0, tx - W % tx)), mode="reflect")  # numpy 'reflect' == REFLECT_101
    tw, th = ext.shape[1] // tx, ext.shape[0] // ty
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip_limit = 0
    if clip > 0:
        clip_limit = max(1, int(clip * area / 256.0))
    luts = np.zeros((ty, tx, 256), dtype=np.uint8)
    for j in range(ty):
        for i in range(tx):
            h = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].reshape(-1), minlength=256).astype(np.int64)
            if clip_limit > 0:
                clipped = int(np.maximum(h - clip_limit, 0).sum())
                h = np.minimum(h, clip_limit)
                batch, residual = clipped // 256, clipped % 256
                h += batch
                if residual:
                    step = max(256 // residual, 1)
                    idx = np.arange(0, 256, step)[:residual]
                    h[idx] += 1
            cs = np.cumsum(h).astype(np.float32)
            luts[j, i] = np.clip(np.rint(cs * lut_scale), 0, 255).astype(np.uint8)
    one, half = np.float32(1.0), np.float32(0.5)

    def axis(n, tsize, ntiles):
        inv = one / np.float32(tsize)
        f = np.arange(n, dtype=np.float32) * inv - half
        t1 = np.floor(f).astype(np.int64)
        a = (f - t1.astype(np.float32)).astype(np.float32)
        a1 = (one - a).astype(np.float32)
        return np.maximum(t1, 0), np.minimum(t1 + 1, ntiles - 1), a, a1

    x1, x2, xa, xa1 = axis(W, tw, tx)
    y1, y2, ya, ya1 = axis(H, th, ty)
    v = y.astype(np.int64)
    f32 = np.float32
    L = luts.astype(f32)
    Y1, Y2 = y1[:, None], y2[:, None]
    X1, X2 = x1[None, :], x2[None, :]
    top = (L[Y1, X1, v] * xa1[None, :]).astype(f32) + (L[Y1, X2, v] * xa[None, :]).astype(f32)
    bot = (L[Y2, X1, v] * xa1[None, :]).astype(f32) + (L[Y2, X2, v] * xa[None, :]).astype(f32)
    res = (top.astype(f32) * ya1[:, None]).astype(f32) + (bot.astype(f32) * ya[:, None]).astype(f32)
    return np.clip(np.rint(res.astype(f32)), 0, 255).astype(np.uint8)


def _descale14(v):
    return (v + 8192) >> 14


def np_bgr2ycc(bgr: np.ndarray, mode=COLOR_YUV) -> np.ndarray:
    B, G, R = (bgr[..., i].astype(np.int64) for i in range(3))
    Y = _descale14(1868 * B + 9617 * G + 4899 * R)
    if mode == COLOR_YUV:
        c1 = _descale14((B - Y) * 8061 + (128 << 14))
        c2 = _descale14((R - Y) * 14369 + (128 << 14))
    else:
        c1 = _descale14((R - Y) * 11682 + (128 << 14))
        c2 = _descale14((B - Y) * 9241 + (128 << 14))
    return np.clip(np.stack([Y, c1, c2], axis=-1), 0, 255).astype(np.uint8)


def np_ycc2bgr(ycc: np.ndarray, mode=COLOR_YUV) -> np.ndarray:
    Y, c1, c2 = (ycc[..., i].astype(np.int64) for i in range(3))
    if mode == COLOR_YUV:
        U, V = c1 - 128, c2 - 128
        B = Y + _descale14(U * 33292)
        G = Y + _descale14(U * -6472 + V * -9519)
        R = Y + _descale14(V * 18678)
    else:
        Cr, Cb = c1 - 128, c2 - 128
        B = Y + _descale14(Cb * 29049)
        G = Y + _descale14(Cb * -5636 + Cr * -11698)
        R = Y + _descale14(Cr * 22987)
    return np.clip(np.stack([B, G, R], axis=-1), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# cv2: the functions the reference itself calls
# --------------------------------------------------------------------------------------------
def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def cv2_nv12_equalize_hist(nv12: np.ndarray, W, H, out: np.ndarray, uv_mode=UV_COPY) -> np.ndarray:
    """Zero-copy frame body of nextimprovement.cpp:159-168 (UV memcpy, then equalizeHist on views)."""
    import cv2
    ysz = W * H
    if uv_mode == UV_COPY:
        out[ysz:] = nv12[ysz:]
    elif uv_mode == UV_GRAY128:
        out[ysz:] = 128
    cv2.equalizeHist(nv12[:ysz].reshape(H, W), out[:ysz].reshape(H, W))
    return out


def cv2_nv12_clahe(nv12: np.ndarray, W, H, out: np.ndarray, clahe=None, clip=2.0, tx=8, ty=8,
                   uv_mode=UV_COPY) -> np.ndarray:
    """Frame body of clahevideo.cpp:178-201 (CLAHE object created once, :497)."""
    import cv2
    if clahe is None:
        clahe = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty))
    ysz = W * H
    if uv_mode == UV_COPY:
        out[ysz:] = nv12[ysz:]
    elif uv_mode == UV_GRAY128:
        out[ysz:] = 128
    clahe.apply(nv12[:ysz].reshape(H, W), out[:ysz].reshape(H, W))
    return out


def cv2_color_equalize(bgr: np.ndarray, mode=COLOR_YUV, use_clahe=False, clip=2.0, tx=8, ty=8) -> np.ndarray:
    """singlecolor.cpp:39-66 / clahe1frame.cpp:83-102."""
    import cv2
    fwd, inv = ((cv2.COLOR_BGR2YUV, cv2.COLOR_YUV2BGR) if mode == COLOR_YUV
                else (cv2.COLOR_BGR2YCrCb, cv2.COLOR_YCrCb2BGR))
    ycc = cv2.cvtColor(bgr, fwd)
    ch = list(cv2.split(ycc))
    if use_clahe:
        ch[0] = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(ch[0])
    else:
        ch[0] = cv2.equalizeHist(ch[0])
    return cv2.cvtColor(cv2.merge(ch), inv)
