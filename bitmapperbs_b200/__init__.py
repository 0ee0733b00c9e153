"""bitmapperbs_b200: B200-native seed-and-verify path of BitMapperBS.

The product is the CUDA library `libbmbs_gpu.so` (C ABI in include/bmbs.h) and the
host mapper `_build/bmbs`; this package is the thin Python host mirror used by the
tests and bench.py (ctypes over the same C ABI).  There is no CPU fallback: loading
fails loudly when the CUDA library has not been built.
"""
from .capi import (BmbsError, Index, Batch, Refiner, Params, ReadResult, Cand, lib_path, load_library,  # noqa: F401
                   NONE, EXACT_UNIQUE, MULTI_EXACT, ONE_MISMATCH, VERIFY)
