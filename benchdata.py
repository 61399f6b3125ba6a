"""Workload preparation for bench.py: synthetic cryo blocks "as found on disk".

A pg_cryogen relation on disk holds blocks that the reference's compression.c wrote
through the system liblz4 / libzstd (compression.c:70-72, :102-104).  To benchmark
*decompression* of such a relation we therefore need compressed inputs produced by
those libraries.  This module calls the system libraries directly (ctypes on
liblz4.so.1 / libzstd.so.1) for that single purpose -- input preparation outside any
timed region.  It is not the product path (libcryogpu.so never calls a CPU codec) and
it is not the oracle (nothing here is used as a checker).
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from pg_cryogen_b200 import blockgen as bg

CRYO_BLCKSZ = 1 << 20
_lz4 = None
_zstd = None


def _libs():
    global _lz4, _zstd
    if _lz4 is None:
        _lz4 = C.CDLL("liblz4.so.1")
        _lz4.LZ4_compress_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lz4.LZ4_compress_fast.restype = C.c_int
        _zstd = C.CDLL("libzstd.so.1")
        _zstd.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        _zstd.ZSTD_compress.restype = C.c_size_t
    return _lz4, _zstd


def library_compress(method: int, level_or_accel: int, block: np.ndarray, scratch: np.ndarray) -> np.ndarray:
    """One block through LZ4_compress_fast / ZSTD_compress, exactly as compression.c calls them."""
    lz4, zstd = _libs()
    if method == 0:
        n = lz4.LZ4_compress_fast(block.ctypes.data, scratch.ctypes.data, block.size, scratch.size,
                                  level_or_accel)
    else:
        n = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, block.ctypes.data, block.size,
                               level_or_accel)
    if n <= 0 or n > scratch.size:
        raise RuntimeError("library compression failed")
    return scratch[:n].copy()


def build_table(nrows: int, kind: str, payload: str, method: int, level_or_accel: int,
                first_block: int = 0, count: int | None = None, keep_plain: int = 32,
                threads: int = 8, block_seed_offset: int = 0):
    """Generate blocks [first_block, first_block+count) of an nrows-row table and compress
    each with the system library.  Returns (list of compressed arrays, plaintext of the
    first `keep_plain` blocks for spot checks)."""
    total = bg.table_block_count(nrows, kind)
    cnt = total - first_block if count is None else count
    per = bg.KINDS[kind][0]
    chunks: list = [None] * cnt
    plain = np.empty((min(keep_plain, cnt), CRYO_BLCKSZ), dtype=np.uint8)

    def work(rng):
        lo, hi = rng
        blk = np.empty(CRYO_BLCKSZ, dtype=np.uint8)
        scratch = np.empty(CRYO_BLCKSZ + (CRYO_BLCKSZ >> 7) + 4096, dtype=np.uint8)
        for i in range(lo, hi):
            b = first_block + i
            n = min(per, nrows - b * per)
            bg.make_block(kind, payload, b + block_seed_offset, ntuples=n, out=blk)
            if i < plain.shape[0]:
                plain[i] = blk
            chunks[i] = library_compress(method, level_or_accel, blk, scratch)

    threads = max(1, min(threads, cnt))
    ranges = [(cnt * t // threads, cnt * (t + 1) // threads) for t in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, ranges))
    return chunks, plain
