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


def _lay_out(nv, planes_y, planes_uv, W, H, lay, size, fill):
    """Place a packed NV12 frame into a buffer with the given plane offsets / strides."""
    buf = np.full(size, fill, np.uint8)
    for r in range(H):
        buf[lay.offset[0] + r * lay.stride[0]: lay.offset[0] + r * lay.stride[0] + W] = planes_y[r]
    for r in range(H // 2):
        buf[lay.offset[1] + r * lay.stride[1]: lay.offset[1] + r * lay.stride[1] + W] = planes_uv[r]
    return buf


@pytest.mark.gpu
def test_gpu_gstvideometa_layouts(nv, oracle):
    """nv12eq_*_meta: planes at arbitrary offsets with padded rows, different layouts for
    the input and the output buffer; payload bit-exact with the oracle, every padding byte untouched."""
    W, H = 322, 202
    packed = oracle.c_synth_nv12(W, H, 2026, 4)
    y, uv = packed[:W * H].reshape(H, W), packed[W * H:].reshape(H // 2, W)
    il = nv.Layout(y_offset=64, uv_offset=64 + 384 * H + 128, y_stride=384, uv_stride=352)
    ol = nv.Layout(y_offset=0, uv_offset=336 * H + 32, y_stride=336, uv_stride=400)
    in_size = il.offset[1] + il.stride[1] * (H // 2) + 17
    out_size = ol.offset[1] + ol.stride[1] * (H // 2) + 5
    src = _lay_out(nv, y, uv, W, H, il, in_size, 0xAB)
    with nv.Context(0, 512, 512, 1) as ctx:
        for op in ("eq", "clahe"):
            for uv_mode, ouv in ((nv.UV_COPY, oracle.UV_COPY), (nv.UV_GRAY128, oracle.UV_GRAY128), (nv.UV_SKIP, oracle.UV_SKIP)):
                want = (oracle.c_nv12_equalize_hist(packed, W, H, uv_mode=ouv, out=np.full_like(packed, 0xCD)) if op == "eq" else
                        oracle.c_nv12_clahe(packed, W, H, 3.0, 4, 5, uv_mode=ouv, out=np.full_like(packed, 0xCD)))
                expect = _lay_out(nv, want[:W * H].reshape(H, W), want[W * H:].reshape(H // 2, W), W, H, ol, out_size, 0xCD)
                out = np.full(out_size, 0xCD, np.uint8)
                if op == "eq":
                    ctx.equalize_hist_meta(src, W, H, il, out, ol, uv_mode)
                else:
                    ctx.clahe_meta(src, W, H, 3.0, (4, 5), il, out, ol, uv_mode)
                assert np.array_equal(out, expect), (op, uv_mode)
        # packed default (NULL layouts) == the plain entry point; in place with one layout
        assert np.array_equal(ctx.equalize_hist_meta(packed, W, H), oracle.c_nv12_equalize_hist(packed, W, H))
        buf = src.copy()
        ctx.equalize_hist_meta(buf, W, H, il, buf, il, nv.UV_COPY)
        want = oracle.c_nv12_equalize_hist(packed, W, H)
        assert np.array_equal(buf, _lay_out(nv, want[:W * H].reshape(H, W), want[W * H:].reshape(H // 2, W), W, H, il, in_size, 0xAB))
        # error behaviour
        bad = nv.Layout(0, W * H, W - 1, W)
        assert ctx.equalize_hist_meta(packed, W, H, bad, raw_status=True) == nv.ERR_INVALID_ARGUMENT
        far = nv.Layout(0, packed.size, W, W)
        assert ctx.equalize_hist_meta(packed, W, H, far, raw_status=True) == nv.ERR_SHORT_BUFFER


# ------------------------------------------------------------------------------------------------------------
# 16-bit CLAHE (CV_16UC1 / P010), SURVEY.md section 8f rank 3
# ------------------------------------------------------------------------------------------------------------
import sys  # noqa: E402
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases_ext import CLAHE16_PARAMS, plane16  # noqa: E402  (deterministic input generators, no cv2)


def test_oracle_clahe16_matches_cv2_golden(oracle):
    for rec in GOLD["clahe16"]:
        if rec["W"] * rec["H"] > 1920 * 1080:
            continue
        y = plane16(rec["W"], rec["H"], rec["kind"], rec["seed"])
        assert sha(y) == rec["in"]
        for key, digest in rec["out"].items():
            clip, tx, ty = key.split(":")
            assert sha(oracle.c_clahe16(y, float(clip), int(tx), int(ty))) == digest, (rec["W"], rec["H"], key)
    assert np.array_equal(oracle.c_clahe16(FIX["clahe16_34x18_in"], 2.0, 4, 3), FIX["clahe16_34x18_2.0_4_3"])


@pytest.mark.gpu
def test_gpu_clahe16_golden_and_p010_frames(nv, oracle):
    with nv.Context(0, 3840, 2160, 1) as ctx:
        for rec in GOLD["clahe16"]:
            y = plane16(rec["W"], rec["H"], rec["kind"], rec["seed"])
            for key, digest in rec["out"].items():
                clip, tx, ty = key.split(":")
                assert sha(ctx.clahe16(y, float(clip), (int(tx), int(ty)))) == digest, (rec["W"], rec["H"], key)
        # strided plane (a view into a wider array)
        wide = np.random.default_rng(3).integers(0, 65536, (50, 90), dtype=np.uint16)
        view = wide[:, 5:5 + 70]
        out = np.zeros_like(wide)
        ctx.clahe16(view, 2.0, (4, 4), out=out[:, 5:5 + 70])
        assert np.array_equal(out[:, 5:75], oracle.c_clahe16(np.ascontiguousarray(view), 2.0, 4, 4))
        assert not out[:, :5].any() and not out[:, 75:].any()
        # P010 frames: Y via the 16-bit CLAHE, chroma copied / neutral / untouched; padded rows stay untouched
        W, H, S = 322, 200, 2 * 322 + 12
        rng = np.random.default_rng(8)
        y = plane16(W, H, "p010", 31)
        uv = (rng.integers(0, 1024, (H // 2, W)).astype(np.uint16) << 6)
        frame = np.full(S * (H + H // 2), 0xEE, np.uint8)
        rows = frame.reshape(H + H // 2, S)
        rows[:H, :2 * W] = y.view(np.uint8).reshape(H, 2 * W)
        rows[H:, :2 * W] = uv.view(np.uint8).reshape(H // 2, 2 * W)
        want_y = oracle.c_clahe16(y, 2.0, 8, 8)
        for uv_mode in (nv.UV_COPY, nv.UV_GRAY128, nv.UV_SKIP):
            out = ctx.p010_clahe(frame, W, H, 2.0, (8, 8), stride=S, uv_mode=uv_mode, out=np.full_like(frame, 0x11))
            orow = out.reshape(H + H // 2, S)
            assert np.array_equal(orow[:H, :2 * W].copy().view(np.uint16).reshape(H, W), want_y), uv_mode
            got_uv = orow[H:, :2 * W].copy().view(np.uint16).reshape(H // 2, W)
            if uv_mode == nv.UV_COPY:
                assert np.array_equal(got_uv, uv)
            elif uv_mode == nv.UV_GRAY128:
                assert (got_uv == 0x8000).all()
            else:
                assert (orow[H:, :2 * W] == 0x11).all()
            assert (orow[:, 2 * W:] == 0x11).all()
        assert ctx.p010_clahe(frame[:100], W, H, raw_status=True, stride=S) == nv.ERR_SHORT_BUFFER
        assert ctx.p010_clahe(frame, W, H, raw_status=True, stride=W) == nv.ERR_INVALID_ARGUMENT


@pytest.mark.gpu
def test_gpu_clahe16_device_batch(nv, oracle):
    import torch
    W, H, n = 640, 360, 20          # 20 planes x 64 tiles: two workspace passes (16 + 4)
    planes = np.stack([plane16(W, H, "p010", 40 + k) for k in range(n)])
    with nv.Context(0, W, H, 1) as ctx:
        d_in = torch.from_numpy(planes.view(np.int16)).cuda()
        d_out = torch.zeros_like(d_in)
        ctx.clahe16_device(d_in, d_out, n, W * H, W, H, 2.0, (8, 8), stream=torch.cuda.current_stream())
        ctx.clahe16_device(d_in, d_out, n, W * H, W, H, 2.0, (8, 8), stream=torch.cuda.current_stream())  # workspace is self-cleaning
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint16)
        for k in (0, 7, 15, 16, 19):
            assert np.array_equal(got[k], oracle.c_clahe16(planes[k], 2.0, 8, 8)), k


@pytest.mark.gpu
def test_gpu_clahe16_bit_depths_in_one_batch(nv, oracle):
    """The cell tables are indexed at v >> z, z = number of low bits that are zero in every pixel of the plane (found on the
    fly, per plane).  One batch mixes planes of every kind: 10-, 12-, 14- and 16-bit content, a plane with a single odd
    pixel (z collapses to 0), constant planes (z = 12, and all-zero: z = 16) -- each must equal OpenCV's 65536-bin result."""
    import torch
    W, H = 200, 120
    rng = np.random.default_rng(77)
    planes = [
        (rng.integers(0, 1024, (H, W)) << 6).astype(np.uint16),
        (rng.integers(0, 4096, (H, W)) << 4).astype(np.uint16),
        (rng.integers(0, 16384, (H, W)) << 2).astype(np.uint16),
        rng.integers(0, 65536, (H, W)).astype(np.uint16),
        (rng.integers(0, 1024, (H, W)) << 6).astype(np.uint16),
        np.full((H, W), 4096, np.uint16),
        np.zeros((H, W), np.uint16),
        (rng.integers(0, 2, (H, W)) << 15).astype(np.uint16),
    ]
    planes[4][H - 1, W - 1] |= 1                      # one odd pixel in the last corner
    batch = np.stack(planes)
    n = len(planes)
    with nv.Context(0, W, H, 1) as ctx:
        d_in = torch.from_numpy(batch.view(np.int16)).cuda()
        for clip, tiles in ((2.0, (8, 8)), (40.0, (3, 5)), (0.0, (1, 1))):
            d_out = torch.zeros_like(d_in)
            ctx.clahe16_device(d_in, d_out, n, W * H, W, H, clip, tiles, stream=torch.cuda.current_stream())
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(np.uint16)
            for k in range(n):
                assert np.array_equal(got[k], oracle.c_clahe16(planes[k], clip, tiles[0], tiles[1])), (k, clip, tiles)


@pytest.mark.gpu
def test_gpu_clahe16_large_planes_mixed_content(nv, oracle):
    """Large planes and tiles (the 4K / 1080p regime of the bench) with 10-bit, full-range and 14-bit content in one launch.
    Odd sizes exercise the padded tiles, the partial row batches and the last column block."""
    import torch
    for (W, H, tiles) in ((1920, 1080, (8, 8)), (1366, 770, (4, 3)), (1000, 523, (2, 1)), (2048, 1100, (7, 8))):
        planes = np.stack([plane16(W, H, "p010", 5), plane16(W, H, "full", 6), plane16(W, H, "p010", 7) >> 2 << 2])
        with nv.Context(0, W, H, 1) as ctx:
            d_in = torch.from_numpy(planes.view(np.int16)).cuda()
            d_out = torch.zeros_like(d_in)
            ctx.clahe16_device(d_in, d_out, 3, W * H, W, H, 3.0, tiles, stream=torch.cuda.current_stream())
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(np.uint16)
            for k in range(3):
                assert np.array_equal(got[k], oracle.c_clahe16(planes[k], 3.0, tiles[0], tiles[1])), (W, H, tiles, k)
