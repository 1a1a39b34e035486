"""CPU tests: the oracle (C restatement + NumPy witness) against the committed cv2-generated golden vectors,
and against live cv2 where it imports.  This is what 'parity pinned' means for this repo."""
import hashlib
import os

import numpy as np
import pytest

from cases import CLAHE_PARAMS, DIST_KINDS, DIST_SIZES, dist_image


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


# SURVEY.md Appendix B known answers (sha1[:16]): W,H,frame -> in Y, in UV, eq(Y), clahe(2.0,8x8)(Y)
APPENDIX_B = {
    (1920, 1080, 0): ("93760852d7a01411", "7925795f58dc1043", "c312b848e15cea35", "52da04b89f65a651"),
    (1920, 1080, 1): ("779530bd35b1df9b", "44bdb5d9905ca29a", "63b2badbe0c0ba86", "f91c80ea539d1be7"),
    (3840, 2160, 0): ("d941f5cfce893d6e", "992401af5e1ae4f2", "17b22cb119487643", "322d2df6af715f64"),
    (3840, 2160, 1): ("3dd6541c21166162", "75e58615a5ea0f51", "15c7a08ff8552efd", "b07c79265d5e2203"),
    (1280, 720, 0): ("0f7def91f34a73e0", "32df3d5f64fe915f", "cc0d3c8d2a76d41a", "3991ead95ccf92df"),
    (1280, 720, 1): ("a93fa82d3b6f2afe", "6c209949e130d004", "962da164a9ac632f", "d7931e3fc5dec31f"),
    (1918, 1078, 0): ("60bb6e59856c233c", "33386c099ff2c726", "8bc4d00060966788", "94244b7b36a793a6"),
    (1918, 1078, 1): ("cc739f626cd4576f", "97e57ce8cf165724", "d5e80a69c23a119e", "e644193c5c057a7e"),
}


@pytest.mark.parametrize("key", sorted(APPENDIX_B))
def test_appendix_b_known_answers(oracle, key):
    W, H, frame = key
    nv = oracle.c_synth_nv12(W, H, 2026, frame)
    y, uv = nv[:W * H].reshape(H, W), nv[W * H:]
    got = (sha(y)[:16], sha(uv)[:16], sha(oracle.c_equalize_hist(y))[:16], sha(oracle.c_clahe(y, 2.0, 8, 8))[:16])
    assert got == APPENDIX_B[key]


def test_synth_generator_two_witnesses(oracle):
    for (W, H, f) in [(64, 48, 0), (1918, 1078, 3), (34, 18, 1), (16, 2, 0)]:
        assert np.array_equal(oracle.c_synth_nv12(W, H, 2026, f), oracle.np_synth_nv12(W, H, 2026, f))


def test_synth_stride_layout(oracle):
    W, H, S = 100, 36, 128
    a = oracle.c_synth_nv12(W, H, 2026, 0)
    b = oracle.c_synth_nv12(W, H, 2026, 0, stride=S)
    assert np.array_equal(b.reshape(H + H // 2, S)[:, :W], a.reshape(H + H // 2, W))
    assert not b.reshape(H + H // 2, S)[:, W:].any()


def test_golden_synth_digests(oracle, golden):
    for rec in golden["synth"]:
        W, H = rec["W"], rec["H"]
        nv = oracle.c_synth_nv12(W, H, rec["seed"], rec["frame"])
        y = nv[:W * H].reshape(H, W)
        assert sha(y) == rec["in_y"] and sha(nv[W * H:]) == rec["in_uv"]
        assert sha(oracle.c_equalize_hist(y)) == rec["eq"], (W, H)
        for key, digest in rec["clahe"].items():
            clip, tx, ty = key.split(":")
            assert sha(oracle.c_clahe(y, float(clip), int(tx), int(ty))) == digest, (W, H, key)


def test_golden_dist_digests(oracle, golden):
    n = 0
    for rec in golden["dist"]:
        y = dist_image(rec["kind"], rec["W"], rec["H"], rec["seed"])
        assert sha(y) == rec["in"], "input generator drifted"
        assert sha(oracle.c_equalize_hist(y)) == rec["eq"], rec
        assert sha(oracle.np_equalize_hist(y)) == rec["eq"], rec
        for key, digest in rec["clahe"].items():
            clip, tx, ty = key.split(":")
            assert sha(oracle.c_clahe(y, float(clip), int(tx), int(ty))) == digest, (rec, key)
            if rec["W"] * rec["H"] <= 128 * 96:
                assert sha(oracle.np_clahe(y, float(clip), int(tx), int(ty))) == digest, (rec, key)
            n += 1
    assert n > 400


def test_raw_fixtures(oracle, fixtures):
    names = [k[:-3] for k in fixtures if k.endswith("_in") and not k.startswith("color")]
    assert names
    for base in names:
        y = fixtures[base + "_in"]
        assert np.array_equal(oracle.c_equalize_hist(y), fixtures[base + "_eq"]), base
        for k in fixtures:
            if k.startswith(base + "_clahe_"):
                clip, tx, ty = k[len(base) + 7:].split("_")
                assert np.array_equal(oracle.c_clahe(y, float(clip), int(tx), int(ty)), fixtures[k]), k


def test_color_golden(oracle, golden, fixtures):
    for rec in golden["color"]:
        W, H = rec["W"], rec["H"]
        if W * H > 1920 * 1080:
            continue  # 4K covered on the GPU side; keep the CPU suite short
        bgr = oracle.c_synth_bgr(W, H, 0)
        assert sha(bgr) == rec["in"]
        for name, mode in (("yuv", oracle.COLOR_YUV), ("ycrcb", oracle.COLOR_YCRCB)):
            assert sha(oracle.c_bgr2ycc(bgr, mode)) == rec[f"{name}_fwd"]
            assert sha(oracle.c_color_equalize(bgr, mode)) == rec[f"{name}_eq"]
            assert sha(oracle.c_color_equalize(bgr, mode, True, 3.0, 4, 4)) == rec[f"{name}_clahe_3.0_4_4"]
    bgr = fixtures["color_31x9_in"]
    assert np.array_equal(oracle.c_color_equalize(bgr, oracle.COLOR_YUV), fixtures["color_31x9_yuv_eq"])
    assert np.array_equal(oracle.np_ycc2bgr(oracle.np_bgr2ycc(bgr, 1), 1), oracle.c_ycc2bgr(oracle.c_bgr2ycc(bgr, 1), 1))


def test_color_full_cube(oracle, golden):
    cube = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([(cube & 255), (cube >> 8) & 255, (cube >> 16) & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    g = golden["cube"]
    assert sha(oracle.c_bgr2ycc(bgr, oracle.COLOR_YUV)) == g["bgr2yuv"]
    assert sha(oracle.c_ycc2bgr(bgr, oracle.COLOR_YUV)) == g["yuv2bgr"]
    assert sha(oracle.c_bgr2ycc(bgr, oracle.COLOR_YCRCB)) == g["bgr2ycrcb"]
    assert sha(oracle.c_ycc2bgr(bgr, oracle.COLOR_YCRCB)) == g["ycrcb2bgr"]


def test_nv12_frame_forms(oracle):
    W, H, S = 70, 38, 96
    nv = oracle.c_synth_nv12(W, H, 2026, 5, stride=S)
    rows = nv.reshape(H + H // 2, S)
    eq = oracle.c_nv12_equalize_hist(nv, W, H, stride=S, uv_mode=oracle.UV_COPY).reshape(H + H // 2, S)
    assert np.array_equal(eq[:H, :W], oracle.c_equalize_hist(np.ascontiguousarray(rows[:H, :W])))
    assert np.array_equal(eq[H:, :W], rows[H:, :W])
    assert not eq[:, W:].any()  # padding bytes are never written
    cl = oracle.c_nv12_clahe(nv, W, H, 2.0, 4, 4, stride=S, uv_mode=oracle.UV_GRAY128).reshape(H + H // 2, S)
    assert np.array_equal(cl[:H, :W], oracle.c_clahe(np.ascontiguousarray(rows[:H, :W]), 2.0, 4, 4))
    assert (cl[H:, :W] == 128).all()
    sk = np.full_like(nv, 7)
    oracle.c_nv12_equalize_hist(nv, W, H, stride=S, uv_mode=oracle.UV_SKIP, out=sk)
    assert (sk.reshape(H + H // 2, S)[H:] == 7).all()


def test_batch_matches_single(oracle):
    W, H, n = 64, 48, 5
    frames = np.stack([oracle.c_synth_nv12(W, H, 2026, k) for k in range(n)])
    eq = oracle.c_nv12_batch("equalize", frames, W, H, threads=2)
    cl = oracle.c_nv12_batch("clahe", frames, W, H, clip=2.0, tx=8, ty=8, threads=2)
    for k in range(n):
        assert np.array_equal(eq[k], oracle.c_nv12_equalize_hist(frames[k], W, H))
        assert np.array_equal(cl[k], oracle.c_nv12_clahe(frames[k], W, H))


def test_live_cv2_when_available(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for it in range(40):
        W, H = int(rng.integers(1, 260)), int(rng.integers(1, 180))
        tx, ty = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        clip = float(rng.choice([0.0, 0.01, 0.5, 1, 2, 3, 4, 40]))
        y = dist_image(DIST_KINDS[it % len(DIST_KINDS)], W, H, 5000 + it)
        assert np.array_equal(cv2.equalizeHist(y), oracle.c_equalize_hist(y))
        try:
            ref = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(y)
        except cv2.error:
            continue
        assert np.array_equal(ref, oracle.c_clahe(y, clip, tx, ty)), (W, H, tx, ty, clip)
    W, H = 320, 180
    nv = oracle.c_synth_nv12(W, H, 2026, 0)
    out = np.zeros_like(nv)
    oracle.cv2_nv12_clahe(nv, W, H, out, clip=2.0, tx=8, ty=8, uv_mode=oracle.UV_GRAY128)
    assert np.array_equal(out, oracle.c_nv12_clahe(nv, W, H, uv_mode=oracle.UV_GRAY128))


@pytest.mark.skipif(not os.path.exists("/root/reference/hun.png"), reason="reference image only in the build container")
def test_hun_png(oracle, golden):
    cv2 = pytest.importorskip("cv2")
    img = cv2.imread("/root/reference/hun.png")
    y = oracle.c_bgr2ycc(img, oracle.COLOR_YUV)[..., 0].copy()
    g = golden["hun"]
    assert list(y.shape) == g["shape"] and sha(y) == g["y"]
    assert sha(oracle.c_equalize_hist(y)) == g["eq"] == "63953e54e66afddaa5331b9cfdfad04e52bd8c6a"
    assert sha(oracle.c_clahe(y, 2.0, 8, 8)) == g["clahe_2.0_8_8"] == "9e872669104a99642348c1e0b1c10476fa69ad38"
    assert sha(oracle.c_color_equalize(img, oracle.COLOR_YUV)) == g["color_yuv_eq"]
