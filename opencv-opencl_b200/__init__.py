"""nv12eq -- B200-native NV12 histogram equalization / CLAHE (host-side Python mirror of the reference call shape).

The reference (kimkimhun3/OpenCV-OpenCL) has no plugin registry; its operator interface for this path is the per-frame
body of the worker (nextimprovement.cpp:128-170, clahevideo.cpp:178-201): an NV12 buffer plus width/height goes in,
an NV12 buffer comes out, with ``cv::equalizeHist`` or ``cv::createCLAHE(clipLimit, Size(t, t))->apply`` run on the
Y view and the chroma plane copied or greyed.  This module keeps those names and argument meanings
(``equalizeHist``, ``createCLAHE(clipLimit=, tileGridSize=)``, ``.apply``, ``setClipLimit``, ``setTilesGridSize``)
over the C-ABI of ``libnv12eq.so`` (``include/nv12eq.h``).  Everything here is a thin ctypes layer: all arithmetic
runs in the hand-written sm_100a CUDA kernels under ``csrc/``.  There is no CPU fallback -- if the shared library is
missing, or no B200 is visible, calls raise.

Import note: the directory name contains a hyphen, so use ``importlib.import_module("opencv-opencl_b200")`` or the
``opencv_opencl_b200`` shim module at the repo root.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NV12EQ_LIB") or os.path.join(_HERE, "libnv12eq.so")  # NV12EQ_LIB: experimental builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nv12eq.h")

# enums of include/nv12eq.h
(OK, ERR_INVALID_ARGUMENT, ERR_SHORT_BUFFER, ERR_CUDA, ERR_NO_DEVICE, ERR_OUT_OF_MEMORY, ERR_BAD_SLOT, ERR_TOO_LARGE, ERR_DROPPED,
 ERR_EMPTY) = range(10)
UV_COPY, UV_GRAY128, UV_SKIP = 0, 1, 2
COLOR_YUV, COLOR_YCRCB = 0, 1
OP_EQUALIZE, OP_CLAHE = 0, 1
FULL_BLOCK, FULL_DROP_NEWEST, FULL_DROP_OLDEST = 0, 1, 2

_u8p = ctypes.POINTER(ctypes.c_uint8)
_c_int, _c_sz, _c_dbl, _c_vp, _c_u32 = ctypes.c_int, ctypes.c_size_t, ctypes.c_double, ctypes.c_void_p, ctypes.c_uint32


class Nv12eqError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"nv12eq status {status}: {message}")
        self.status = status


class Counters(ctypes.Structure):
    _fields_ = [("frames", ctypes.c_uint64), ("bytes_in", ctypes.c_uint64), ("bytes_out", ctypes.c_uint64),
                ("errors", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("busy_us", ctypes.c_uint64)]


class StreamConfig(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int), ("width", ctypes.c_int), ("height", ctypes.c_int), ("stride", ctypes.c_int),
                ("uv_mode", ctypes.c_int), ("clip_limit", ctypes.c_double), ("tiles_x", ctypes.c_int), ("tiles_y", ctypes.c_int),
                ("depth", ctypes.c_int), ("full_policy", ctypes.c_int)]


class Layout(ctypes.Structure):
    """nv12eq_layout: GstVideoMeta-style plane offsets and strides (plane 0 = Y, plane 1 = UV)."""
    _fields_ = [("offset", ctypes.c_size_t * 2), ("stride", ctypes.c_int * 2)]

    def __init__(self, y_offset=0, uv_offset=0, y_stride=0, uv_stride=0):
        super().__init__((ctypes.c_size_t * 2)(y_offset, uv_offset), (ctypes.c_int * 2)(y_stride, uv_stride))


class StreamStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint64) for k in ("pushed", "delivered", "dropped_backpressure", "in_flight", "max_in_flight",
                                               "latency_us_sum", "latency_us_max")]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ into libnv12eq.so for sm_100a with nvcc (in-tree, so the .so travels with the repo)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))] + [HEADER_PATH]
    if os.environ.get("NV12EQ_LIB"):
        return LIB_PATH  # an explicitly selected experimental build is used as is
    def stale():
        return (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale():
        # several ranks of one torchrun job may get here at once: one builds, the others wait and then find it fresh
        import fcntl
        with open(os.path.join(csrc, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    out = subprocess.run(["make", "-C", csrc] + (["-B"] if force else []), capture_output=True, text=True)
                    if verbose or out.returncode:
                        print(out.stdout + out.stderr)
                    if out.returncode:
                        raise RuntimeError("building libnv12eq.so failed (see output above)")
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


TMA_LIB_PATH = os.path.join(_HERE, "libnv12eq_tma.so")


def build_variants(verbose: bool = False) -> str:
    """Compile the measured-and-rejected TMA variant of the library (CLAHE tile rows staged by cp.async.bulk.tensor + mbarriers, colour
    rounds as bulk copies) next to the shipped one; tests/test_variants.py loads it through NV12EQ_LIB in a subprocess."""
    csrc = os.path.join(_HERE, "csrc")
    out = subprocess.run(["make", "-C", csrc, "variants"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout + out.stderr)
    if out.returncode:
        raise RuntimeError("building libnv12eq_tma.so failed (see output above)")
    return TMA_LIB_PATH


_SIGNATURES = {
    # name: (restype, argtypes)
    "nv12eq_version": (_c_int, []),
    "nv12eq_status_string": (ctypes.c_char_p, [_c_int]),
    "nv12eq_last_error_string": (ctypes.c_char_p, [_c_vp]),
    "nv12eq_create": (_c_int, [_c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_vp)]),
    "nv12eq_destroy": (None, [_c_vp]),
    "nv12eq_get_counters": (_c_int, [_c_vp, ctypes.POINTER(Counters)]),
    "nv12eq_set_tuning": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_host_alloc": (_c_int, [_c_sz, ctypes.POINTER(_c_vp)]),
    "nv12eq_host_free": (_c_int, [_c_vp]),
    "nv12eq_equalize_hist": (_c_int, [_c_vp, _c_vp, _c_sz, _c_vp, _c_sz, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_clahe": (_c_int, [_c_vp, _c_vp, _c_sz, _c_vp, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int]),
    "nv12eq_equalize_hist_batch": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_clahe_batch": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int]),
    "nv12eq_submit_equalize_hist": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_submit_clahe": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int]),
    "nv12eq_wait": (_c_int, [_c_vp, _c_int]),
    "nv12eq_query": (_c_int, [_c_vp, _c_int]),
    "nv12eq_equalize_hist_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_clahe_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_sync": (_c_int, [_c_vp]),
    "nv12eq_hist_device": (_c_int, [_c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "nv12eq_equalize_apply_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_vp, ctypes.c_int64, _c_vp]),
    "nv12eq_clahe_band_luts_device": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, ctypes.c_double, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "nv12eq_clahe_band_apply_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "nv12eq_color_equalize": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_color_equalize_batch": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int]),
    "nv12eq_color_clahe": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int]),
    "nv12eq_color_equalize_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_color_clahe_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_vp]),
    "nv12eq_equalize_hist_meta": (_c_int, [_c_vp, _c_vp, _c_sz, ctypes.POINTER(Layout), _c_vp, _c_sz, ctypes.POINTER(Layout), _c_int, _c_int,
                                           _c_int]),
    "nv12eq_clahe_meta": (_c_int, [_c_vp, _c_vp, _c_sz, ctypes.POINTER(Layout), _c_vp, _c_sz, ctypes.POINTER(Layout), _c_int, _c_int, _c_dbl,
                                   _c_int, _c_int, _c_int]),
    "nv12eq_clahe16_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_vp]),
    "nv12eq_clahe16": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int]),
    "nv12eq_p010_clahe": (_c_int, [_c_vp, _c_vp, _c_sz, _c_vp, _c_sz, _c_int, _c_int, _c_int, _c_dbl, _c_int, _c_int, _c_int]),
    "nv12eq_bgr_to_i420": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_sz]),
    "nv12eq_bgr_to_i420_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_sz, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_nv12_to_bgr": (_c_int, [_c_vp, _c_vp, _c_sz, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_int]),
    "nv12eq_bgr_to_nv12": (_c_int, [_c_vp, _c_vp, _c_sz, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_int]),
    "nv12eq_nv12_to_bgr_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_sz, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_bgr_to_nv12_device": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_sz, _c_sz, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "nv12eq_stream_open": (_c_int, [_c_vp, ctypes.POINTER(StreamConfig), ctypes.POINTER(_c_vp)]),
    "nv12eq_stream_push": (_c_int, [_c_vp, _c_vp, _c_sz, ctypes.POINTER(ctypes.c_uint64)]),
    "nv12eq_stream_pop": (_c_int, [_c_vp, _c_vp, _c_sz, ctypes.POINTER(ctypes.c_uint64), _c_int]),
    "nv12eq_stream_get_stats": (_c_int, [_c_vp, ctypes.POINTER(StreamStats)]),
    "nv12eq_stream_close": (None, [_c_vp]),
    "nv12eq_synth_nv12_device": (_c_int, [_c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_u32, _c_u32, _c_vp]),
    "nv12eq_synth_bgr_device": (_c_int, [_c_vp, _c_vp, _c_int, _c_sz, _c_int, _c_int, _c_int, _c_u32, _c_vp]),
}

_lib: Optional[ctypes.CDLL] = None


def load_library() -> ctypes.CDLL:
    """dlopen libnv12eq.so and bind every symbol include/nv12eq.h declares.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() / make -C "
                              f"{os.path.join(_HERE, 'csrc')}.  nv12eq has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def nv12_frame_bytes(width: int, height: int, stride: Optional[int] = None) -> int:
    stride = width if stride is None else stride
    return stride * (height + height // 2)


def _ptr(buf) -> int:
    """Address of a NumPy array, torch tensor (host or device), ctypes buffer or raw integer."""
    if buf is None:
        return 0
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        return int(buf.data_ptr())
    return ctypes.addressof(buf)


def _nbytes(buf) -> int:
    if isinstance(buf, np.ndarray):
        return buf.nbytes
    if hasattr(buf, "numel"):
        return int(buf.numel() * buf.element_size())
    return ctypes.sizeof(buf)


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy
CUDA_STREAM_PER_THREAD = 2  # cudaStreamPerThread


def _stream_ptr(stream) -> int:
    """None -> 0 (the context's own stream).  A torch.cuda.Stream whose handle is 0 is the legacy default stream; the
    C-ABI reserves NULL for 'own stream', so it is passed as cudaStreamLegacy."""
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    h = int(stream.cuda_stream)  # torch.cuda.Stream
    return h if h != 0 else CUDA_STREAM_LEGACY


class PinnedBuffer:
    """Page-locked host memory from nv12eq_host_alloc, viewed as a uint8 NumPy array (``.array``)."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        p = _c_vp()
        st = self._lib.nv12eq_host_alloc(nbytes, ctypes.byref(p))
        if st != OK:
            raise Nv12eqError(st, f"nv12eq_host_alloc({nbytes}) failed")
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.nv12eq_host_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One device context (the reference's per-worker OpenCL context, OpenCLequalHist.cpp:71-81,142-161).

    Thread-compatible: use one Context per calling thread.
    """

    def __init__(self, device: int = 0, max_width: int = 4096, max_height: int = 2304, slots: int = 2):
        self._lib = load_library()
        h = _c_vp()
        st = self._lib.nv12eq_create(device, max_width, max_height, slots, ctypes.byref(h))
        if st != OK:
            msg = self._lib.nv12eq_last_error_string(None).decode()
            raise Nv12eqError(st, msg or self._lib.nv12eq_status_string(st).decode())
        self._h = h
        self.device = device
        self.slots = slots

    # -- plumbing ---------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.nv12eq_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != OK:
            raise Nv12eqError(st, self._lib.nv12eq_last_error_string(self._h).decode()
                              or self._lib.nv12eq_status_string(st).decode())

    def status(self, st: int) -> str:
        return self._lib.nv12eq_status_string(st).decode()

    def last_error(self) -> str:
        return self._lib.nv12eq_last_error_string(self._h).decode()

    def counters(self) -> dict:
        c = Counters()
        self._check(self._lib.nv12eq_get_counters(self._h, ctypes.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def set_tuning(self, chunks_per_frame: int = 0, lag_frames: int = 0, ctas_per_sm: int = 0, schedule: int = 0):
        self._check(self._lib.nv12eq_set_tuning(self._h, chunks_per_frame, lag_frames, ctas_per_sm, schedule))

    def sync(self):
        self._check(self._lib.nv12eq_sync(self._h))

    # -- host frame in / frame out ----------------------------------------------------------------------
    def equalize_hist(self, nv12, width: int, height: int, stride: Optional[int] = None, uv_mode: int = UV_COPY,
                      out=None, raw_status: bool = False):
        """nextimprovement.cpp:159-168: out = NV12 frame with equalizeHist(Y) and the chroma plane per uv_mode."""
        stride = width if stride is None else stride
        out = np.empty_like(nv12) if out is None else out
        st = self._lib.nv12eq_equalize_hist(self._h, _ptr(nv12), _nbytes(nv12), _ptr(out), _nbytes(out), width, height,
                                            stride, uv_mode)
        if raw_status:
            return st
        self._check(st)
        return out

    def clahe(self, nv12, width: int, height: int, clip_limit: float = 2.0, tiles: Tuple[int, int] = (8, 8),
              stride: Optional[int] = None, uv_mode: int = UV_COPY, out=None, raw_status: bool = False):
        """clahevideo.cpp:178-201: out = NV12 frame with CLAHE(Y); tiles = (tilesX, tilesY) as in cv::Size."""
        stride = width if stride is None else stride
        out = np.empty_like(nv12) if out is None else out
        st = self._lib.nv12eq_clahe(self._h, _ptr(nv12), _nbytes(nv12), _ptr(out), _nbytes(out), width, height, stride,
                                    float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode)
        if raw_status:
            return st
        self._check(st)
        return out

    def equalize_hist_meta(self, buf, width, height, in_layout=None, out=None, out_layout=None, uv_mode=UV_COPY, raw_status=False):
        """Frame whose planes sit at GstVideoMeta offsets/strides (``Layout``); None = packed planes."""
        out = np.empty_like(buf) if out is None else out
        st = self._lib.nv12eq_equalize_hist_meta(self._h, _ptr(buf), _nbytes(buf), ctypes.byref(in_layout) if in_layout else None,
                                                 _ptr(out), _nbytes(out), ctypes.byref(out_layout) if out_layout else None, width,
                                                 height, uv_mode)
        if raw_status:
            return st
        self._check(st)
        return out

    def clahe_meta(self, buf, width, height, clip_limit=2.0, tiles=(8, 8), in_layout=None, out=None, out_layout=None,
                   uv_mode=UV_COPY, raw_status=False):
        out = np.empty_like(buf) if out is None else out
        st = self._lib.nv12eq_clahe_meta(self._h, _ptr(buf), _nbytes(buf), ctypes.byref(in_layout) if in_layout else None, _ptr(out),
                                         _nbytes(out), ctypes.byref(out_layout) if out_layout else None, width, height,
                                         float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode)
        if raw_status:
            return st
        self._check(st)
        return out

    def equalize_hist_batch(self, frames, width, height, stride=None, uv_mode=UV_COPY, out=None, n_frames=None,
                            frame_pitch=None):
        stride = width if stride is None else stride
        n, pitch = self._batch_shape(frames, n_frames, frame_pitch)
        out = np.empty_like(frames) if out is None else out
        self._check(self._lib.nv12eq_equalize_hist_batch(self._h, _ptr(frames), _ptr(out), n, pitch, width, height, stride,
                                                         uv_mode))
        return out

    def clahe_batch(self, frames, width, height, clip_limit=2.0, tiles=(8, 8), stride=None, uv_mode=UV_COPY, out=None,
                    n_frames=None, frame_pitch=None):
        stride = width if stride is None else stride
        n, pitch = self._batch_shape(frames, n_frames, frame_pitch)
        out = np.empty_like(frames) if out is None else out
        self._check(self._lib.nv12eq_clahe_batch(self._h, _ptr(frames), _ptr(out), n, pitch, width, height, stride,
                                                 float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode))
        return out

    @staticmethod
    def _batch_shape(frames, n_frames, frame_pitch):
        if n_frames is not None:
            return int(n_frames), int(frame_pitch)
        assert frames.ndim == 2, "batch must be (n_frames, frame_bytes)"
        return int(frames.shape[0]), int(frames.shape[1] * frames.itemsize if isinstance(frames, np.ndarray)
                                         else frames.stride(0) * frames.element_size())

    # -- asynchronous slots -----------------------------------------------------------------------------
    def submit_equalize_hist(self, slot, frames, out, n_frames, frame_pitch, width, height, stride=None, uv_mode=UV_COPY):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_submit_equalize_hist(self._h, slot, _ptr(frames), _ptr(out), n_frames, frame_pitch,
                                                          width, height, stride, uv_mode))

    def submit_clahe(self, slot, frames, out, n_frames, frame_pitch, width, height, clip_limit=2.0, tiles=(8, 8),
                     stride=None, uv_mode=UV_COPY):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_submit_clahe(self._h, slot, _ptr(frames), _ptr(out), n_frames, frame_pitch, width,
                                                  height, stride, float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode))

    def wait(self, slot: int):
        self._check(self._lib.nv12eq_wait(self._h, slot))

    def query(self, slot: int) -> bool:
        st = self._lib.nv12eq_query(self._h, slot)
        if st == OK:
            return True
        if st == ERR_BAD_SLOT:
            return False
        self._check(st)
        return False

    # -- device-resident forms (torch tensors or raw device addresses) -----------------------------------
    def equalize_hist_device(self, d_in, d_out, n_frames, frame_pitch, width, height, stride=None, uv_mode=UV_COPY,
                             stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_equalize_hist_device(self._h, _ptr(d_in), _ptr(d_out), n_frames, frame_pitch, width,
                                                          height, stride, uv_mode, _stream_ptr(stream)))

    def clahe_device(self, d_in, d_out, n_frames, frame_pitch, width, height, clip_limit=2.0, tiles=(8, 8), stride=None,
                     uv_mode=UV_COPY, stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_clahe_device(self._h, _ptr(d_in), _ptr(d_out), n_frames, frame_pitch, width, height,
                                                  stride, float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode,
                                                  _stream_ptr(stream)))

    def hist_device(self, d_y, n_planes, plane_pitch, width, height, d_hist, stride=None, stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_hist_device(self._h, _ptr(d_y), n_planes, plane_pitch, width, height, stride,
                                                 _ptr(d_hist), _stream_ptr(stream)))

    def equalize_apply_device(self, d_y_in, d_y_out, n_planes, plane_pitch, width, height, d_hist, total_pixels,
                              stride=None, stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_equalize_apply_device(self._h, _ptr(d_y_in), _ptr(d_y_out), n_planes, plane_pitch,
                                                           width, height, stride, _ptr(d_hist), int(total_pixels),
                                                           _stream_ptr(stream)))

    def clahe_band_luts_device(self, d_y_band, width, full_height, clip_limit, tiles, first_tile_row, band_tiles_y, d_luts,
                               stride=None, stream=None):
        """Tile LUTs of tile rows [first_tile_row, first_tile_row + band_tiles_y) of one frame (spatial split, stage 1)."""
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_clahe_band_luts_device(self._h, _ptr(d_y_band), width, full_height, stride, float(clip_limit),
                                                            int(tiles[0]), int(tiles[1]), first_tile_row, band_tiles_y, _ptr(d_luts),
                                                            _stream_ptr(stream)))

    def clahe_band_apply_device(self, d_y_band, d_out_band, width, full_height, tiles, first_tile_row, band_tiles_y, d_luts_halo,
                                stride=None, stream=None):
        """Interpolation of the band from its LUT grid with one halo tile row on either side (spatial split, stage 2)."""
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_clahe_band_apply_device(self._h, _ptr(d_y_band), _ptr(d_out_band), width, full_height, stride,
                                                             int(tiles[0]), int(tiles[1]), first_tile_row, band_tiles_y,
                                                             _ptr(d_luts_halo), _stream_ptr(stream)))

    # -- colour path ------------------------------------------------------------------------------------
    def color_equalize(self, bgr: np.ndarray, color_mode: int = COLOR_YUV, out=None) -> np.ndarray:
        """singlecolor.cpp:39-66 on a (H, W, 3) uint8 BGR image."""
        h, w, _ = bgr.shape
        out = np.empty_like(bgr) if out is None else out
        self._check(self._lib.nv12eq_color_equalize(self._h, _ptr(bgr), _ptr(out), w, h, bgr.strides[0], color_mode))
        return out

    def color_equalize_batch(self, frames: np.ndarray, color_mode: int = COLOR_YUV, out=None) -> np.ndarray:
        """(n, H, W, 3) uint8 BGR frames through the colour path, pipelined over the context's slots."""
        n, h, w, _ = frames.shape
        out = np.empty_like(frames) if out is None else out
        self._check(self._lib.nv12eq_color_equalize_batch(self._h, _ptr(frames), _ptr(out), n, frames.strides[0], w, h,
                                                          frames.strides[1], color_mode))
        return out

    def color_clahe(self, bgr: np.ndarray, clip_limit=3.0, tiles=(4, 4), color_mode: int = COLOR_YUV, out=None):
        """clahe1frame.cpp:83-102 (defaults clip 3.0, tile 4: clahe1frame.cpp:55-56)."""
        h, w, _ = bgr.shape
        out = np.empty_like(bgr) if out is None else out
        self._check(self._lib.nv12eq_color_clahe(self._h, _ptr(bgr), _ptr(out), w, h, bgr.strides[0], color_mode,
                                                 float(clip_limit), int(tiles[0]), int(tiles[1])))
        return out

    def color_equalize_device(self, d_in, d_out, n_frames, frame_pitch, width, height, stride=None, color_mode=COLOR_YUV,
                              stream=None):
        stride = 3 * width if stride is None else stride
        self._check(self._lib.nv12eq_color_equalize_device(self._h, _ptr(d_in), _ptr(d_out), n_frames, frame_pitch, width,
                                                           height, stride, color_mode, _stream_ptr(stream)))

    def color_clahe_device(self, d_in, d_out, n_frames, frame_pitch, width, height, clip_limit=3.0, tiles=(4, 4),
                           stride=None, color_mode=COLOR_YUV, stream=None):
        stride = 3 * width if stride is None else stride
        self._check(self._lib.nv12eq_color_clahe_device(self._h, _ptr(d_in), _ptr(d_out), n_frames, frame_pitch, width,
                                                        height, stride, color_mode, float(clip_limit), int(tiles[0]),
                                                        int(tiles[1]), _stream_ptr(stream)))

    # -- 16-bit CLAHE (CV_16UC1 / P010) ----------------------------------------------------------------
    def clahe16(self, plane: np.ndarray, clip_limit: float = 2.0, tiles: Tuple[int, int] = (8, 8), out=None) -> np.ndarray:
        """``cv2.createCLAHE(clip, tiles).apply`` on a (H, W) uint16 plane (65536-bin path)."""
        assert plane.dtype == np.uint16 and plane.ndim == 2
        h, w = plane.shape
        out = np.empty_like(plane) if out is None else out
        self._check(self._lib.nv12eq_clahe16(self._h, _ptr(plane), _ptr(out), w, h, plane.strides[0] // 2, float(clip_limit),
                                             int(tiles[0]), int(tiles[1])))
        return out

    def clahe16_device(self, d_in, d_out, n_planes, plane_pitch, width, height, clip_limit=2.0, tiles=(8, 8), stride=None,
                       stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_clahe16_device(self._h, _ptr(d_in), _ptr(d_out), n_planes, plane_pitch, width, height, stride,
                                                    float(clip_limit), int(tiles[0]), int(tiles[1]), _stream_ptr(stream)))

    def p010_clahe(self, frame, width, height, clip_limit=2.0, tiles=(8, 8), stride=None, uv_mode=UV_COPY, out=None,
                   raw_status=False):
        """P010 frame (bytes): Y through the 16-bit CLAHE, chroma per uv_mode.  stride in bytes (default 2*width)."""
        stride = 2 * width if stride is None else stride
        out = np.empty_like(frame) if out is None else out
        st = self._lib.nv12eq_p010_clahe(self._h, _ptr(frame), _nbytes(frame), _ptr(out), _nbytes(out), width, height, stride,
                                         float(clip_limit), int(tiles[0]), int(tiles[1]), uv_mode)
        if raw_status:
            return st
        self._check(st)
        return out

    # -- BGR -> I420 adapter ---------------------------------------------------------------------------
    def bgr_to_i420(self, bgr: np.ndarray, out=None) -> np.ndarray:
        """``cv2.cvtColor(bgr, COLOR_BGR2YUV_I420)`` (1frameMeasure.cpp:32): (H, W, 3) uint8 -> (H*3/2, W) planar."""
        h, w, _ = bgr.shape
        out = np.empty((h * 3 // 2, w), np.uint8) if out is None else out
        self._check(self._lib.nv12eq_bgr_to_i420(self._h, _ptr(bgr), w, h, bgr.strides[0], _ptr(out), _nbytes(out)))
        return out

    def bgr_to_i420_device(self, d_bgr, d_out, n_frames, bgr_pitch, out_pitch, width, height, stride=None, stream=None):
        stride = 3 * width if stride is None else stride
        self._check(self._lib.nv12eq_bgr_to_i420_device(self._h, _ptr(d_bgr), _ptr(d_out), n_frames, bgr_pitch, out_pitch, width,
                                                        height, stride, _stream_ptr(stream)))

    # -- NV12 <-> BGR adapters -------------------------------------------------------------------------
    def nv12_to_bgr(self, nv12: np.ndarray, width: int, height: int, stride: Optional[int] = None, out=None) -> np.ndarray:
        """``cv2.cvtColor(nv12, COLOR_YUV2BGR_NV12)``: flat NV12 frame -> (H, W, 3) BGR (display side of the NV12 path)."""
        stride = width if stride is None else stride
        out = np.empty((height, width, 3), np.uint8) if out is None else out
        span = out.strides[0] * (height - 1) + 3 * width   # bytes from the first to the last pixel (rows may be strided views)
        self._check(self._lib.nv12eq_nv12_to_bgr(self._h, _ptr(nv12), _nbytes(nv12), width, height, stride, _ptr(out), span,
                                                 out.strides[0]))
        return out

    def bgr_to_nv12(self, bgr: np.ndarray, stride: Optional[int] = None, out=None) -> np.ndarray:
        """BGR -> flat NV12 frame (COLOR_BGR2YUV_I420 arithmetic, chroma interleaved U first): the input of the NV12 operators."""
        h, w, _ = bgr.shape
        stride = w if stride is None else stride
        out = np.empty(stride * (h + h // 2), np.uint8) if out is None else out
        self._check(self._lib.nv12eq_bgr_to_nv12(self._h, _ptr(bgr), bgr.strides[0] * (h - 1) + 3 * w, w, h, bgr.strides[0], _ptr(out),
                                                 _nbytes(out), stride))
        return out

    def nv12_to_bgr_device(self, d_nv12, d_bgr, n_frames, nv12_pitch, bgr_pitch, width, height, stride=None, bgr_stride=None, stream=None):
        stride = width if stride is None else stride
        bgr_stride = 3 * width if bgr_stride is None else bgr_stride
        self._check(self._lib.nv12eq_nv12_to_bgr_device(self._h, _ptr(d_nv12), _ptr(d_bgr), n_frames, nv12_pitch, bgr_pitch, width, height,
                                                        stride, bgr_stride, _stream_ptr(stream)))

    def bgr_to_nv12_device(self, d_bgr, d_nv12, n_frames, bgr_pitch, nv12_pitch, width, height, bgr_stride=None, stride=None, stream=None):
        stride = width if stride is None else stride
        bgr_stride = 3 * width if bgr_stride is None else bgr_stride
        self._check(self._lib.nv12eq_bgr_to_nv12_device(self._h, _ptr(d_bgr), _ptr(d_nv12), n_frames, bgr_pitch, nv12_pitch, width, height,
                                                        bgr_stride, stride, _stream_ptr(stream)))

    # -- synthetic inputs -------------------------------------------------------------------------------
    def synth_nv12_device(self, d_out, n_frames, frame_pitch, width, height, stride=None, seed=2026, first_frame=0,
                          stream=None):
        stride = width if stride is None else stride
        self._check(self._lib.nv12eq_synth_nv12_device(self._h, _ptr(d_out), n_frames, frame_pitch, width, height, stride,
                                                       seed, first_frame, _stream_ptr(stream)))

    def synth_bgr_device(self, d_out, n_frames, frame_pitch, width, height, stride=None, first_frame=0, stream=None):
        stride = 3 * width if stride is None else stride
        self._check(self._lib.nv12eq_synth_bgr_device(self._h, _ptr(d_out), n_frames, frame_pitch, width, height, stride,
                                                      first_frame, _stream_ptr(stream)))


class Stream:
    """Ordered, back-pressured frame stream over one context (``nv12eq_stream_*``): the reference's worker queue
    (OpenCVequalHist.cpp:71-98,397-402) with in-order delivery.  ``push`` returns the frame's sequence number or None
    when the back-pressure policy dropped it; ``pop`` returns ``(seq, frame)`` or None when nothing is ready."""

    def __init__(self, ctx: Context, width: int, height: int, op: int = OP_EQUALIZE, stride: Optional[int] = None,
                 uv_mode: int = UV_COPY, clip_limit: float = 2.0, tiles: Tuple[int, int] = (8, 8), depth: int = 8,
                 full_policy: int = FULL_BLOCK):
        self._ctx, self._lib = ctx, ctx._lib
        stride = width if stride is None else stride
        self.frame_bytes = nv12_frame_bytes(width, height, stride)
        cfg = StreamConfig(op, width, height, stride, uv_mode, float(clip_limit), int(tiles[0]), int(tiles[1]), depth, full_policy)
        h = _c_vp()
        ctx._check(self._lib.nv12eq_stream_open(ctx._h, ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h

    def push(self, nv12) -> Optional[int]:
        seq = ctypes.c_uint64()
        st = self._lib.nv12eq_stream_push(self._h, _ptr(nv12), _nbytes(nv12), ctypes.byref(seq))
        if st == ERR_DROPPED:
            return None
        self._ctx._check(st)
        return int(seq.value)

    def pop(self, out=None, block: bool = True):
        out = np.empty(self.frame_bytes, np.uint8) if out is None else out
        seq = ctypes.c_uint64()
        st = self._lib.nv12eq_stream_pop(self._h, _ptr(out), _nbytes(out), ctypes.byref(seq), 1 if block else 0)
        if st == ERR_EMPTY:
            return None
        self._ctx._check(st)
        return int(seq.value), out

    def stats(self) -> dict:
        s = StreamStats()
        self._ctx._check(self._lib.nv12eq_stream_get_stats(self._h, ctypes.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in StreamStats._fields_}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nv12eq_stream_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------------
# Reference-shaped operator interface (names and argument meaning of the OpenCV calls the reference makes)
# ------------------------------------------------------------------------------------------------------------
_default_ctx: Optional[Context] = None


def default_slots() -> int:
    """Host lanes a context gets by default for batch calls: three, so that an upload, a kernel and a download are always
    queued on different lanes (two lanes leave the copy engines idle while the host re-arms a lane)."""
    return 3


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def equalizeHist(nv12: np.ndarray, width: int, height: int, dst: Optional[np.ndarray] = None, uv_mode: int = UV_COPY,
                 ctx: Optional[Context] = None) -> np.ndarray:
    """``cv::equalizeHist(y_plane_in, y_plane_out)`` on the Y view of an NV12 buffer + chroma passthrough
    (nextimprovement.cpp:159-168)."""
    return (ctx or default_context()).equalize_hist(nv12, width, height, uv_mode=uv_mode, out=dst)


class CLAHE:
    """``cv::Ptr<cv::CLAHE>`` as the reference uses it (clahevideo.cpp:184-195): created once, ``apply`` per frame."""

    def __init__(self, clipLimit: float = 40.0, tileGridSize: Sequence[int] = (8, 8), ctx: Optional[Context] = None):
        self._clip = float(clipLimit)
        self._tiles = (int(tileGridSize[0]), int(tileGridSize[1]))
        self._ctx = ctx

    def setClipLimit(self, clipLimit: float):
        self._clip = float(clipLimit)

    def getClipLimit(self) -> float:
        return self._clip

    def setTilesGridSize(self, tileGridSize: Sequence[int]):
        self._tiles = (int(tileGridSize[0]), int(tileGridSize[1]))

    def getTilesGridSize(self) -> Tuple[int, int]:
        return self._tiles

    def apply(self, nv12: np.ndarray, width: int, height: int, dst: Optional[np.ndarray] = None,
              uv_mode: int = UV_COPY) -> np.ndarray:
        return (self._ctx or default_context()).clahe(nv12, width, height, self._clip, self._tiles, uv_mode=uv_mode, out=dst)


def createCLAHE(clipLimit: float = 40.0, tileGridSize: Sequence[int] = (8, 8), ctx: Optional[Context] = None) -> CLAHE:
    """Same defaults as ``cv::createCLAHE`` (40.0, 8x8); the reference passes 2.0 / 8 (clahevideo.cpp:384-385)."""
    return CLAHE(clipLimit, tileGridSize, ctx)


from . import sharding  # noqa: E402  (multi-GPU partitioning helpers: shard_range, Reassembler, FrameShardedStream, ...)
