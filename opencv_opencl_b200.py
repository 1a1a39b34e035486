"""Import shim: the package directory is ``opencv-opencl_b200`` (not a valid identifier), so
``import opencv_opencl_b200 as nv12eq`` re-exports it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("opencv-opencl_b200")
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
