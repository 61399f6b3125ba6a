"""-m gpu: cryogpu_{de,}compress_host_multi -- the batch cut into contiguous block ranges, one per
context, no collective (blocks are independent: compression.c takes one block, no
dictionary).  Run with one context always and with two when the box has two GPUs; an odd block
count and mixed per-block methods (storage.h:64) so that the ranges differ in size and content."""
import ctypes as C

import numpy as np
import pytest

from pg_cryogen_b200 import COMP_LZ4, COMP_ZSTD, CRYO_BLCKSZ, CryoGPU, compress_bound
from pg_cryogen_b200 import blockgen as bg

pytestmark = pytest.mark.gpu


def _contexts(k):
    return [CryoGPU(i) for i in range(k)]


def _ctx_array(ctxs):
    return (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])


def _ngpus(lib):
    return int(lib.cryogpu_device_count())


@pytest.mark.parametrize("nctx", [1, 2])
def test_decompress_host_multi_mixed_methods(gpu, oracle_ref, nctx):
    if nctx > _ngpus(gpu.lib):
        pytest.skip("one GPU on this box")
    kinds = [("S", "hex"), ("M", "lowcard"), ("D", "hex"), ("S", "lowcard"), ("M", "hex")]
    n = 7
    blocks = np.stack([bg.make_block(*kinds[i % len(kinds)], 70 + i) for i in range(n)])
    methods = np.array([i % 2 for i in range(n)], dtype=np.int32)
    z = oracle_ref.compress(COMP_ZSTD, 1, blocks, nthreads=4)[0]
    l = oracle_ref.compress(COMP_LZ4, 1, blocks, nthreads=4)[0]
    chunks = [np.ascontiguousarray(z[i] if methods[i] else l[i]) for i in range(n)]
    chunks[5] = chunks[5][:-9].copy()                   # one bad block must not fail the batch
    sizes = np.array([c.size for c in chunks], dtype=np.uint32)
    out = np.full((n, CRYO_BLCKSZ), 0x77, dtype=np.uint8)
    srcp = (C.c_void_p * n)(*[c.ctypes.data for c in chunks])
    dstp = (C.c_void_p * n)(*[out[i].ctypes.data for i in range(n)])
    osz = np.zeros(n, dtype=np.uint32)
    st = np.full(n, -1, dtype=np.int32)
    ctxs = _contexts(nctx)
    try:
        rc = gpu.lib.cryogpu_decompress_host_multi(_ctx_array(ctxs), nctx, n, methods.ctypes.data, srcp,
                                                   sizes.ctypes.data, dstp, CRYO_BLCKSZ,
                                                   osz.ctypes.data, st.ctypes.data)
        assert rc == 0, gpu.lib.cryogpu_last_error().decode()
    finally:
        for c in ctxs:
            c.close()
    good = [i for i in range(n) if i != 5]
    assert (st[good] == 0).all() and st[5] != 0, st
    assert (osz[good] == CRYO_BLCKSZ).all()
    assert np.array_equal(out[good], blocks[good])


@pytest.mark.parametrize("nctx", [1, 2])
@pytest.mark.parametrize("method,level", [(COMP_LZ4, 1), (COMP_ZSTD, 1)])
def test_compress_host_multi_roundtrips_through_the_reference(gpu, oracle_ref, nctx, method, level):
    if nctx > _ngpus(gpu.lib):
        pytest.skip("one GPU on this box")
    n = 5
    blocks = np.stack([bg.make_block("SMD"[i % 3], "lowcard" if i & 1 else "hex", 90 + i) for i in range(n)])
    bound = compress_bound(method)
    dst = np.zeros((n, bound), dtype=np.uint8)
    srcp = (C.c_void_p * n)(*[blocks[i].ctypes.data for i in range(n)])
    dstp = (C.c_void_p * n)(*[dst[i].ctypes.data for i in range(n)])
    sz = np.zeros(n, dtype=np.uint32)
    st = np.full(n, -1, dtype=np.int32)
    ctxs = _contexts(nctx)
    try:
        rc = gpu.lib.cryogpu_compress_host_multi(_ctx_array(ctxs), nctx, n, method, level, srcp,
                                                 CRYO_BLCKSZ, dstp, bound, sz.ctypes.data, st.ctypes.data)
        assert rc == 0, gpu.lib.cryogpu_last_error().decode()
    finally:
        for c in ctxs:
            c.close()
    assert (st == 0).all(), st
    comp = [dst[i, : sz[i]].copy() for i in range(n)]
    back, ok, _ = oracle_ref.decompress([method] * n, *oracle_ref.pack(comp))
    assert ok.all() and np.array_equal(back, blocks)


def test_host_multi_rejects_empty_context_list(gpu):
    rc = gpu.lib.cryogpu_decompress_host_multi(None, 0, 1, None, None, None, None, CRYO_BLCKSZ, None, None)
    assert rc != 0
