"""ctypes binding of the page split / gather restatement in oracle/libcryooracle.so (TEST INFRASTRUCTURE
ONLY): oracle/cryo_pages.c follows pg_cryogen.c:689-805 and cache.c:100-176 over a numpy "relation"."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import port

PAGE = 8192
ERR_SUCCESS, ERR_WRONG_STARTING_BLOCK, ERR_EMPTY_BLOCK = 0, 2, 3


def _lib():
    L = port.lib()
    L.cryo_oracle_pages_needed.restype = C.c_uint32
    L.cryo_oracle_pages_needed.argtypes = [C.c_uint64]
    L.cryo_oracle_pages_split.restype = C.c_uint32
    L.cryo_oracle_pages_split.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.cryo_oracle_pages_gather.restype = C.c_int
    L.cryo_oracle_pages_gather.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
    return L


def layout() -> list[int]:
    out = (C.c_uint32 * 12)()
    _lib().cryo_oracle_page_layout(out)
    return list(out)


def pages_needed(size: int) -> int:
    return int(_lib().cryo_oracle_pages_needed(size))


def split(relation: np.ndarray, blknos, comp, method: int, xid: int) -> int:
    """Write the compressed block `comp` into relation[blknos[k]] (pg_cryogen.c:761-805).  Returns npages."""
    comp = np.ascontiguousarray(comp, dtype=np.uint8)
    b = np.ascontiguousarray(blknos, dtype=np.uint32)
    assert len(b) >= pages_needed(comp.size)
    return int(_lib().cryo_oracle_pages_split(relation.ctypes.data, relation.shape[0], b.ctypes.data, comp.ctypes.data,
                                              comp.size, method, xid))


def gather(relation: np.ndarray, block: int, cap: int = (1 << 20) + (1 << 13)):
    """cache.c:100-176 -> (err, method, compressed_size, bytes, chain)"""
    out = np.zeros(cap, dtype=np.uint8)
    method, size, got, nb = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    blocks = np.zeros(relation.shape[0] + 1, dtype=np.uint32)
    err = _lib().cryo_oracle_pages_gather(relation.ctypes.data, relation.shape[0], block, out.ctypes.data, cap,
                                          C.byref(method), C.byref(size), C.byref(got), blocks.ctypes.data, C.byref(nb))
    return int(err), int(method.value), int(size.value), out[: got.value].copy(), [int(x) for x in blocks[: nb.value]]


def block_tuple_stats(block: np.ndarray):
    """(ntuples, tuple_bytes, valid) of one decoded cryo block: the item walk of a sequential scan."""
    L = _lib()
    b = np.ascontiguousarray(block, dtype=np.uint8)
    n, by, ok = C.c_uint32(0), C.c_uint64(0), C.c_int32(0)
    L.cryo_oracle_block_tuple_stats(C.c_void_p(b.ctypes.data), C.c_uint32(b.size), C.byref(n), C.byref(by), C.byref(ok))
    return int(n.value), int(by.value), int(ok.value)
