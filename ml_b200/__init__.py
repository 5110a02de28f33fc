"""B200-native clustering hot path of ML++ (ml::EM, ml::Clustering::KMeans).

Layout:
    csrc/   CUDA kernels (sm_100a, FP64) and the C-ABI of include/mlb200.h  -> lib/libmlb200.so
    host/   the reference's C++ class API and the cppyml pybind module over that C-ABI -> lib/libML.so, host/cppyml
    cabi.py ctypes binding of the C-ABI (tests and bench.py call the device path through it)
There is no CPU fallback: without the built CUDA library or without a device everything fails loudly.
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CPPYML_PATH = os.path.join(ROOT, "host", "cppyml")


def import_cppyml():
    """Imports the cppyml package built under ml_b200/host/cppyml (the reference's Python package name)."""
    if CPPYML_PATH not in sys.path:
        sys.path.insert(0, CPPYML_PATH)
    import cppyml
    return cppyml
