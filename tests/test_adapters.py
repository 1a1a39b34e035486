"""BGR -> I420 adapter (cvtColor COLOR_BGR2YUV_I420, 1frameMeasure.cpp:32): the oracle
against cv2-generated golden vectors on the CPU, and the CUDA path through the C-ABI against both on the GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_ext.json")))
FIX = dict(np.load(os.path.join(HERE, "golden", "fixtures_ext.npz")))


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def corners():
    return np.array([[[b, g, r] for b in (0, 255) for g in (0, 255)] for r in (0, 255)], dtype=np.uint8).reshape(2, 4, 3)


def test_oracle_i420_matches_cv2_golden(oracle):
    for rec in GOLD["i420_synth"]:
        if rec["W"] * rec["H"] > 1920 * 1080:
            continue
        bgr = oracle.c_synth_bgr(rec["W"], rec["H"], 0)
        assert sha(bgr) == rec["in"]
        out = oracle.c_bgr2i420(bgr)
        assert sha(out) == rec["i420"], (rec["W"], rec["H"])
        assert sha(oracle.c_equalize_hist(out[:rec["H"]].copy())) == rec["y_eq"]
    for rec in GOLD["i420_random"]:
        bgr = np.random.default_rng(rec["seed"]).integers(0, 256, (rec["H"], rec["W"], 3), dtype=np.uint8)
        assert sha(bgr) == rec["in"] and sha(oracle.c_bgr2i420(bgr)) == rec["i420"]
    assert sha(oracle.c_bgr2i420(corners())) == GOLD["i420_corners"]
    assert np.array_equal(oracle.c_bgr2i420(FIX["i420_6x4_in"]), FIX["i420_6x4_out"])
    with pytest.raises(ValueError):
        oracle.c_bgr2i420(np.zeros((3, 4, 3), np.uint8))


def test_oracle_i420_live_cv2_when_available(oracle):
    if not oracle.have_cv2():
        pytest.skip("cv2 not importable")
    import cv2
    bgr = np.random.default_rng(99).integers(0, 256, (38, 54, 3), dtype=np.uint8)
    assert np.array_equal(oracle.c_bgr2i420(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420))


@pytest.fixture(scope="module")
def nv():
    import opencv_opencl_b200 as nv12eq
    nv12eq.build()
    return nv12eq


@pytest.mark.gpu
def test_gpu_i420_golden_and_oracle(nv, oracle):
    with nv.Context(0, 3840, 2160, 1) as ctx:
        for rec in GOLD["i420_synth"]:
            bgr = oracle.c_synth_bgr(rec["W"], rec["H"], 0)
            out = ctx.bgr_to_i420(bgr)
            assert sha(out) == rec["i420"], (rec["W"], rec["H"])
        for rec in GOLD["i420_random"]:
            bgr = np.random.default_rng(rec["seed"]).integers(0, 256, (rec["H"], rec["W"], 3), dtype=np.uint8)
            assert sha(ctx.bgr_to_i420(bgr)) == rec["i420"]
        assert sha(ctx.bgr_to_i420(corners())) == GOLD["i420_corners"]
        # strided rows (a view into a wider image) and the 1frameMeasure shape: equalizeHist on the I420 Y plane
        wide = np.random.default_rng(5).integers(0, 256, (40, 70, 3), dtype=np.uint8)
        view = wide[:, 3:3 + 62]
        assert np.array_equal(ctx.bgr_to_i420(view), oracle.c_bgr2i420(np.ascontiguousarray(view)))
        W, H = 640, 360
        bgr = oracle.c_synth_bgr(W, H, 1)
        i420 = ctx.bgr_to_i420(bgr)
        frame = np.concatenate([i420[:H].reshape(-1), np.full(W * H // 2, 128, np.uint8)])   # Y + neutral NV12 chroma
        got = ctx.equalize_hist(frame, W, H, uv_mode=nv.UV_SKIP, out=frame.copy())
        assert np.array_equal(got[:W * H].reshape(H, W), oracle.c_equalize_hist(oracle.c_bgr2i420(bgr)[:H].copy()))
        # error behaviour: odd sizes are rejected like OpenCV does, short output buffers like the frame entry points
        odd = np.zeros((5, 4, 3), np.uint8)
        assert ctx._lib.nv12eq_bgr_to_i420(ctx._h, odd.ctypes.data, 4, 5, 12, odd.ctypes.data, 100) == nv.ERR_INVALID_ARGUMENT
        small = np.zeros(10, np.uint8)
        ok = np.zeros((4, 4, 3), np.uint8)
        assert ctx._lib.nv12eq_bgr_to_i420(ctx._h, ok.ctypes.data, 4, 4, 12, small.ctypes.data, 10) == nv.ERR_SHORT_BUFFER


@pytest.mark.gpu
def test_gpu_i420_device_batch(nv, oracle):
    import torch
    W, H, n = 322, 200, 5
    pitch, opitch = 3 * W * H, W * H * 3 // 2
    with nv.Context(0, W, H, 1) as ctx:
        st = torch.cuda.current_stream()
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros(n * opitch, dtype=torch.uint8, device="cuda")
        ctx.synth_bgr_device(d_in, n, pitch, W, H, first_frame=0, stream=st)
        ctx.bgr_to_i420_device(d_in, d_out, n, pitch, opitch, W, H, stream=st)
        torch.cuda.synchronize()
        for k in range(n):
            want = oracle.c_bgr2i420(oracle.c_synth_bgr(W, H, k))
            assert np.array_equal(d_out[k * opitch:(k + 1) * opitch].cpu().numpy().reshape(H * 3 // 2, W), want), k
