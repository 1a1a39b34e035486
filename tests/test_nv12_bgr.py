"""NV12 <-> BGR adapters: cvtColor(COLOR_YUV2BGR_NV12) for the display side of the NV12 path and the
COLOR_BGR2YUV_I420 arithmetic with interleaved chroma in front of it.  The oracle against cv2-generated golden vectors on the CPU,
the CUDA path through the C-ABI against both on the GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_nv12bgr.json")))
FIX = dict(np.load(os.path.join(HERE, "golden", "fixtures_nv12bgr.npz")))


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_pair(oracle, W, H):
    nv12 = oracle.c_synth_nv12(W, H, 2026, 0).copy()
    nv12[W * H:] = oracle.c_synth_nv12(W, H, 7026, 1)[:W * H // 2]
    return nv12, oracle.c_synth_bgr(W, H, 0)


def random_pair(W, H, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)


def cases(oracle, max_pixels):
    for rec in GOLD["synth"]:
        if rec["W"] * rec["H"] <= max_pixels:
            yield rec, synth_pair(oracle, rec["W"], rec["H"])
    for rec in GOLD["random"]:
        yield rec, random_pair(rec["W"], rec["H"], rec["seed"])


def test_oracle_matches_cv2_golden(oracle):
    for rec, (nv12, bgr) in cases(oracle, 1920 * 1080):
        W, H = rec["W"], rec["H"]
        assert sha(nv12) == rec["nv12_in"] and sha(bgr) == rec["bgr_in"]
        assert sha(oracle.c_nv12_to_bgr(nv12, W, H)) == rec["bgr"], (W, H)
        assert sha(oracle.c_bgr_to_nv12(bgr)) == rec["nv12"], (W, H)
    e = GOLD["extremes"]
    assert sha(oracle.c_nv12_to_bgr(FIX["extremes_nv12"], e["W"], e["H"])) == e["bgr"]
    assert np.array_equal(oracle.c_nv12_to_bgr(FIX["nv12_6x4_in"], 6, 4), FIX["bgr_6x4_out"])
    assert np.array_equal(oracle.c_bgr_to_nv12(FIX["bgr_6x4_in"]), FIX["nv12_6x4_out"])
    with pytest.raises(ValueError):
        oracle.c_bgr_to_nv12(np.zeros((3, 4, 3), np.uint8))


def test_oracle_live_cv2_when_available(oracle):
    cv2 = pytest.importorskip("cv2")
    nv12, bgr = random_pair(130, 70, 5)
    assert np.array_equal(oracle.c_nv12_to_bgr(nv12, 130, 70), cv2.cvtColor(nv12.reshape(105, 130), cv2.COLOR_YUV2BGR_NV12))
    i420 = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420).reshape(-1)
    n = 130 * 70
    want = np.concatenate([i420[:n], np.stack([i420[n:n + n // 4], i420[n + n // 4:]], -1).reshape(-1)])
    assert np.array_equal(oracle.c_bgr_to_nv12(bgr), want)


@pytest.fixture(scope="module")
def nv():
    import opencv_opencl_b200 as nv12eq
    nv12eq.build()
    return nv12eq


@pytest.mark.gpu
def test_gpu_golden_and_oracle(nv, oracle):
    with nv.Context(0, 3840, 2160, 1) as ctx:
        for rec, (nv12, bgr) in cases(oracle, 3840 * 2160):
            W, H = rec["W"], rec["H"]
            assert sha(ctx.nv12_to_bgr(nv12, W, H)) == rec["bgr"], (W, H)
            assert sha(ctx.bgr_to_nv12(bgr)) == rec["nv12"], (W, H)
        e = GOLD["extremes"]
        assert sha(ctx.nv12_to_bgr(FIX["extremes_nv12"], e["W"], e["H"])) == e["bgr"]
        # strided rows on both sides: bytes between the rows stay untouched
        W, H, S, BS = 250, 130, 256, 3 * 250 + 10
        nv12, bgr = random_pair(W, H, 9)
        padded = np.full(S * (H + H // 2), 7, np.uint8)
        padded.reshape(-1, S)[:, :W] = nv12.reshape(-1, W)
        out = np.full((H, BS), 9, np.uint8)
        view = np.lib.stride_tricks.as_strided(out, (H, W, 3), (BS, 3, 1))
        ctx.nv12_to_bgr(padded, W, H, stride=S, out=view)
        assert np.array_equal(view, oracle.c_nv12_to_bgr(nv12, W, H)) and (out[:, 3 * W:] == 9).all()
        src = np.full((H, BS), 3, np.uint8)
        sview = np.lib.stride_tricks.as_strided(src, (H, W, 3), (BS, 3, 1))
        sview[...] = bgr
        got = ctx.bgr_to_nv12(sview, stride=S, out=np.full(S * (H + H // 2), 5, np.uint8))
        assert np.array_equal(got.reshape(-1, S)[:, :W].reshape(-1), oracle.c_bgr_to_nv12(bgr)) and (got.reshape(-1, S)[:, W:] == 5).all()
        # round trip through the operator: BGR -> NV12 -> equalizeHist -> BGR equals the same chain on the CPU
        W, H = 322, 200
        _, bgr = synth_pair(oracle, W, H)
        frame = ctx.bgr_to_nv12(bgr)
        eq = ctx.equalize_hist(frame, W, H)
        assert np.array_equal(ctx.nv12_to_bgr(eq, W, H), oracle.c_nv12_to_bgr(oracle.c_nv12_equalize_hist(oracle.c_bgr_to_nv12(bgr), W, H), W, H))
        # error behaviour
        z = np.zeros(100, np.uint8)
        assert ctx._lib.nv12eq_nv12_to_bgr(ctx._h, z.ctypes.data, 100, 5, 4, 5, z.ctypes.data, 100, 15) == nv.ERR_INVALID_ARGUMENT
        assert ctx._lib.nv12eq_nv12_to_bgr(ctx._h, z.ctypes.data, 10, 4, 4, 4, z.ctypes.data, 100, 12) == nv.ERR_SHORT_BUFFER
        assert ctx._lib.nv12eq_bgr_to_nv12(ctx._h, z.ctypes.data, 100, 4, 4, 12, z.ctypes.data, 10, 4) == nv.ERR_SHORT_BUFFER


@pytest.mark.gpu
def test_gpu_device_batches(nv, oracle):
    import torch
    W, H, n = 322, 200, 5
    bp, npitch = 3 * W * H, W * H * 3 // 2
    with nv.Context(0, W, H, 1) as ctx:
        st = torch.cuda.current_stream()
        d_bgr = torch.empty(n * bp, dtype=torch.uint8, device="cuda")
        d_nv12 = torch.zeros(n * npitch, dtype=torch.uint8, device="cuda")
        d_back = torch.zeros(n * bp, dtype=torch.uint8, device="cuda")
        ctx.synth_bgr_device(d_bgr, n, bp, W, H, first_frame=0, stream=st)
        ctx.bgr_to_nv12_device(d_bgr, d_nv12, n, bp, npitch, W, H, stream=st)
        ctx.nv12_to_bgr_device(d_nv12, d_back, n, npitch, bp, W, H, stream=st)
        torch.cuda.synchronize()
        for k in range(n):
            want = oracle.c_bgr_to_nv12(oracle.c_synth_bgr(W, H, k))
            assert np.array_equal(d_nv12[k * npitch:(k + 1) * npitch].cpu().numpy(), want), k
            assert np.array_equal(d_back[k * bp:(k + 1) * bp].cpu().numpy().reshape(H, W, 3), oracle.c_nv12_to_bgr(want, W, H)), k
