"""ctypes binding of oracle/libcryooracle.so (TEST INFRASTRUCTURE ONLY): the plain-C
restatement of the LZ4 block and zstd frame decoders in oracle/cryo_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libcryooracle.so")
_lib = None


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "sequences", "literal_bytes", "match_bytes", "longest_match", "overlap_match_bytes",
        "max_literal_run", "max_offset", "frames", "window_size", "single_segment",
        "blocks_raw", "blocks_rle", "blocks_compressed",
        "lit_raw", "lit_rle", "lit_huf1", "lit_huf4", "lit_treeless1", "lit_treeless4",
        "huf_direct_weights", "huf_fse_weights",
        "mode_predef", "mode_rle", "mode_fse", "mode_repeat", "rep_offsets")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


SEQ_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32)


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "libcryooracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        L = C.CDLL(_PATH)
        for f in (L.cryo_oracle_lz4_decode, L.cryo_oracle_zstd_decode):
            f.restype = C.c_long
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Stats),
                          C.c_void_p, C.c_void_p]
        L.cryo_oracle_xxh64.restype = C.c_uint64
        L.cryo_oracle_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        _lib = L
    return _lib


def _decode(fn, comp, cap, want_stats, trace):
    c = np.frombuffer(bytes(comp), dtype=np.uint8) if not isinstance(comp, np.ndarray) else \
        np.ascontiguousarray(comp)
    out = np.zeros(cap, dtype=np.uint8)
    st = Stats()
    seqs = []
    cb = SEQ_CB(lambda ctx, ll, ml, off: seqs.append((ll, ml, off))) if trace else None
    r = fn(c.ctypes.data if c.size else None, c.size, out.ctypes.data, cap, C.byref(st),
           C.cast(cb, C.c_void_p) if cb else None, None)
    res = [int(r), out]
    if want_stats:
        res.append(st.as_dict())
    if trace:
        res.append(seqs)
    return tuple(res)


def lz4_decode(comp, cap: int = 1 << 20, stats: bool = False, trace: bool = False):
    """-> (bytes written or negative error, out[cap] [, stats] [, sequences])"""
    return _decode(lib().cryo_oracle_lz4_decode, comp, cap, stats, trace)


def zstd_decode(comp, cap: int = 1 << 20, stats: bool = False, trace: bool = False):
    return _decode(lib().cryo_oracle_zstd_decode, comp, cap, stats, trace)


def xxh64(data, seed: int = 0) -> int:
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    return int(lib().cryo_oracle_xxh64(b.ctypes.data if b.size else None, b.size, seed))
