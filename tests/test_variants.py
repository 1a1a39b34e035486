"""The measured-and-rejected TMA variant of the library stays buildable and bit-exact: libnv12eq_tma.so stages the CLAHE tile rows with
cp.async.bulk.tensor boxes + mbarrier stages and feeds the colour kernel with cp.async.bulk rounds (profiles/r02_clahe_notes.md has the A/B)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tma_lib():
    import opencv_opencl_b200 as nv12eq
    return nv12eq.build_variants()


def test_variant_contains_tma_instructions(tma_lib):
    sass = subprocess.run(["cuobjdump", "-sass", tma_lib], capture_output=True, text=True).stdout
    assert "UTMALDG.3D" in sass          # cp.async.bulk.tensor.3d: tile rows of the CLAHE histogram pass
    assert "UBLKCP" in sass              # cp.async.bulk: colour kernel rounds
    assert "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass   # mbarrier expect_tx / try_wait
    assert "FFMA2" not in sass           # the blend stays unfused


CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, torch
import opencv_opencl_b200 as nv
from oracle import oracle as O
ok = True
for (W, H, n, tiles) in ((1920, 1080, 5, 8), (3840, 2160, 2, 8), (640, 480, 3, 4)):
    pitch = nv.nv12_frame_bytes(W, H)
    with nv.Context(0, W, H, 1) as ctx:
        d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); d_out = torch.zeros_like(d_in)
        st = torch.cuda.current_stream()
        ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
        for _ in range(2):
            ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (tiles, tiles), stream=st)
        torch.cuda.synchronize()
        for k in range(n):
            fr = d_in[k * pitch:(k + 1) * pitch].cpu().numpy()
            ok &= bool(np.array_equal(d_out[k * pitch:(k + 1) * pitch].cpu().numpy(), O.c_nv12_clahe(fr, W, H, 2.0, tiles, tiles)))
        bgr = O.c_synth_bgr(W, H, 1)
        ok &= bool(np.array_equal(ctx.color_equalize(bgr), O.c_color_equalize(bgr, O.COLOR_YUV)))
print("VARIANT_OK" if ok else "VARIANT_MISMATCH")
"""


@pytest.mark.gpu
def test_variant_is_bit_exact_on_the_gpu(tma_lib):
    env = dict(os.environ, NV12EQ_LIB=tma_lib)
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert "VARIANT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
