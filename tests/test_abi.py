"""CPU tests of the drop-in boundary: the shared library builds for sm_100a, loads without a GPU, exports exactly the
symbols include/nv12eq.h declares, fails loudly (no CPU fallback) when no device is present, and the product never
touches the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nv():
    import opencv_opencl_b200 as nv12eq
    nv12eq.build()
    return nv12eq


def header_functions():
    text = open(os.path.join(ROOT, "include", "nv12eq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nv12eq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ("nv12eq_create", "nv12eq_destroy", "nv12eq_equalize_hist", "nv12eq_clahe", "nv12eq_equalize_hist_batch",
                 "nv12eq_clahe_batch", "nv12eq_submit_equalize_hist", "nv12eq_submit_clahe", "nv12eq_wait",
                 "nv12eq_equalize_hist_device", "nv12eq_clahe_device", "nv12eq_color_equalize", "nv12eq_last_error_string",
                 "nv12eq_hist_device", "nv12eq_equalize_apply_device"):
        assert must in fns
    # every declaration cites the reference interface it replaces
    head = open(os.path.join(ROOT, "include", "nv12eq.h")).read()
    for cite in ("nextimprovement.cpp:128-170", "clahevideo.cpp:178-201", "OpenCLequalHist.cpp:349-365", "accel.cpp:36-61",
                 "singlecolor.cpp:39-66"):
        assert cite in head


def test_library_exports_every_declared_symbol(nv):
    lib = ctypes.CDLL(nv.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in nv12eq.h but not exported by libnv12eq.so"
    out = subprocess.run(["nm", "-D", "--defined-only", nv.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(nv12eq_[a-z0-9_]+)\b", out)))
    assert exported == header_functions(), "exported nv12eq_* symbols and header declarations differ"


def test_python_binding_table_matches_header(nv):
    assert sorted(nv._SIGNATURES) == header_functions()
    lib = nv.load_library()
    assert lib.nv12eq_version() == 2   # NV12EQ_VERSION_MAJOR * 100 + NV12EQ_VERSION_MINOR
    assert lib.nv12eq_status_string(nv.ERR_SHORT_BUFFER) == b"buffer too small for the frame"


def test_library_is_sm100a_only(nv):
    out = subprocess.run(["cuobjdump", "-lelf", nv.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_contracted_float_arithmetic(nv):
    """OpenCV's CLAHE blend rounds every product and sum separately; a fused multiply-add anywhere in
    the library would round once instead of twice.  ptxas is known to contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2,
    so the build is checked: no FFMA2 anywhere, and no FFMA at all inside the CLAHE kernels (integer IMAD is fine; the
    scalar FFMAs in equalize_kernel are the Newton steps of the correctly rounded __fdiv_rn(255, total - hist[i0]))."""
    sass = subprocess.run(["cuobjdump", "-sass", nv.LIB_PATH], capture_output=True, text=True).stdout
    assert "clahe_kernel" in sass
    assert not re.findall(r"\bFFMA2\b[^;]*;", sass)
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        if "clahe_kernel" in fn.splitlines()[0] or "clahe16_interp_kernel" in fn.splitlines()[0]:
            fused = re.findall(r"\bFFMA\b[^;]*;", fn)
            assert not fused, fused[:4]
    assert "FMUL2" in sass and "FADD2.FTZ" in sass  # the packed, unfused blend is what was built


def test_no_device_fails_loudly(nv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nv.Nv12eqError) as e:
        nv.Context()
    assert e.value.status == nv.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    lib = nv.load_library()
    assert lib.nv12eq_equalize_hist(None, None, 0, None, 0, 16, 16, 16, 0) == nv.ERR_INVALID_ARGUMENT


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "opencv-opencl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "libnv12eq_oracle", "oracle/", "oracle.py", "import cv2"):
                    assert bad not in text, f"{f} reaches for the CPU checker: {bad}"
    text = open(os.path.join(ROOT, "opencv_opencl_b200.py")).read()
    assert "oracle" not in text and "cv2" not in text


def test_missing_library_is_an_import_error(nv, tmp_path, monkeypatch):
    monkeypatch.setattr(nv._pkg if hasattr(nv, "_pkg") else nv, "LIB_PATH", str(tmp_path / "nope.so"), raising=False)
    import importlib
    pkg = importlib.import_module("opencv-opencl_b200")
    monkeypatch.setattr(pkg, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(pkg, "_lib", None)
    with pytest.raises(ImportError):
        pkg.load_library()


def test_cpp_example_builds_and_fails_loudly_without_a_gpu(nv):
    """examples/worker_demo.cpp is the reference's worker loop with the OpenCV calls replaced by the C-ABI, in C++.  It
    must compile against include/nv12eq.h + libnv12eq.so with plain g++; without a GPU it reports the error and exits 1
    (frames are counted as errors and dropped, like the reference's processing_errors)."""
    ex = os.path.join(ROOT, "examples")
    subprocess.run(["make", "-C", ex, "-B", "-s"], check=True)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the run is covered by tests/test_gpu_stream.py")
    out = subprocess.run([os.path.join(ex, "worker_demo"), "--frames", "3", "--width", "64", "--height", "32"],
                         capture_output=True, text=True)
    assert out.returncode == 1
    assert "no CPU fallback" in out.stderr and "errors" in out.stdout
