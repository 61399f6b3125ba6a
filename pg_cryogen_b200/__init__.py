"""pg_cryogen_b200 -- B200 (sm_100a) block codec for pg_cryogen.

The product is ``libcryogpu.so`` (hand-written CUDA kernels behind the C ABI in
``include/cryogpu.h``) plus ``host/compression.c``, the drop-in for the reference's
compression.c.  This Python package is a thin ctypes binding used by the tests
and the benchmark; it contains no codec logic and no CPU fallback.
"""
from .codec import (COMP_LZ4, COMP_ZSTD, CRYO_BLCKSZ, CryoGPU, CryoGPUError, STATUS_NAMES,
                    compress_bound, lib_path, load_library)

__all__ = ["COMP_LZ4", "COMP_ZSTD", "CRYO_BLCKSZ", "CryoGPU", "CryoGPUError", "STATUS_NAMES",
           "compress_bound", "lib_path", "load_library"]
