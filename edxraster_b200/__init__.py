"""edxraster_b200 — B200-native (sm_100a) implementation of EDXRaster's raster hot path.

The product is the C-ABI shared library built from csrc/ (include/edxraster_c.h) and the C++ host
API above it (include/edxraster/*.h). This Python package is the harness: ctypes bindings,
procedural workloads and the multi-GPU frame farm used by tests/ and bench.py.
"""
__version__ = "0.1.0"
