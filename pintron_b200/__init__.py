"""pintron_b200 — B200-native est-fact hot path (PIntron).

The product is the C host program `pintron_b200/host/est-fact` plus `libpintron_cuda.so` (C ABI in
include/pintron_cuda.h).  This Python package is only the ctypes mirror of that ABI used by tests, bench.py
and __graft_entry__: it never computes anything itself and raises if the CUDA library is missing.
"""
from .binding import (Cuda, Batch, PC_OP, library_path, build_library, LibraryMissing)  # noqa: F401
