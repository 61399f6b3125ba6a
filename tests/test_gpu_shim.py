"""-m gpu: the reference's own API on a B200.

libcryo_compression.so (pg_cryogen_b200/host/compression.c) exports exactly what the reference's
compression.h:13-24 declares.  These tests call cryo_define_compression_gucs / cryo_compress /
cryo_decompress the way pg_cryogen.c:173, :726 and cache.c:178 do, one block per call, and check
them against the reference itself (oracle/_ref = the reference's compression.c on liblz4/libzstd):
GPU-compressed blocks must be restored by the reference, reference-compressed blocks by the shim,
GUC values are read at call time (compression.c:72, :104), malformed input returns false
(compression.c:85-86, :117-118), and an unknown method takes the elog(ERROR) path
(compression.c:137, :157).
"""
import ctypes as C
import os

import numpy as np
import pytest

from pg_cryogen_b200 import COMP_LZ4, COMP_ZSTD, CRYO_BLCKSZ, compress_bound
from pg_cryogen_b200 import blockgen as bg

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    path = os.path.join(HERE, "libshimharness.so")
    assert os.path.exists(path), "tests/libshimharness.so is not built (run __graft_entry__.build())"
    L = C.CDLL(path)
    L.shim_define_gucs.argtypes = [C.POINTER(C.c_int)] * 3
    L.shim_set_gucs.argtypes = [C.c_int] * 3
    L.shim_compress.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                C.c_char_p, C.c_size_t]
    L.shim_decompress.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_char_p, C.c_size_t]
    m, a, l = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    L.shim_define_gucs(C.byref(m), C.byref(a), C.byref(l))
    assert (m.value, a.value, l.value) == (COMP_ZSTD, 1, 1)        # compression.c:24-58 boot values

    class Shim:
        lib = L

        @staticmethod
        def set_gucs(method=COMP_ZSTD, accel=1, level=1):
            L.shim_set_gucs(method, accel, level)

        @staticmethod
        def compress(method, block):
            block = np.ascontiguousarray(block, dtype=np.uint8)
            cap = compress_bound(COMP_LZ4) + 64
            out = np.zeros(cap, dtype=np.uint8)
            sz = C.c_size_t(0)
            msg = C.create_string_buffer(256)
            rc = L.shim_compress(method, block.ctypes.data, out.ctypes.data, cap, C.byref(sz), msg, 256)
            return rc, out[: sz.value].copy(), msg.value.decode()

        @staticmethod
        def decompress(method, comp, fill=0xA5):
            comp = np.ascontiguousarray(comp, dtype=np.uint8)
            out = np.full(CRYO_BLCKSZ, fill, dtype=np.uint8)
            msg = C.create_string_buffer(256)
            rc = L.shim_decompress(method, comp.ctypes.data if comp.size else None, comp.size,
                                   out.ctypes.data, msg, 256)
            return rc, out, msg.value.decode()

    yield Shim
    L.shim_shutdown()


def _blocks():
    return [("S/hex", bg.make_block("S", "hex", 3)), ("M/lowcard", bg.make_block("M", "lowcard", 4)),
            ("D/hex", bg.make_block("D", "hex", 5)), ("D/lowcard", bg.make_block("D", "lowcard", 6)),
            ("regression", bg.regression_block(1, 290))]


@pytest.mark.parametrize("method", [COMP_LZ4, COMP_ZSTD])
def test_shim_compress_is_read_by_the_reference(shim, oracle_ref, method):
    """cryo_compress on the GPU -> the reference's cryo_decompress (compression.c:144-159)."""
    shim.set_gucs()
    for tag, blk in _blocks():
        rc, comp, msg = shim.compress(method, blk)
        assert rc == 0, (tag, rc, msg)
        assert 0 < len(comp) <= compress_bound(method)
        back, ok = oracle_ref.decompress_one(method, comp)
        assert ok and np.array_equal(back, blk), tag


@pytest.mark.parametrize("method", [COMP_LZ4, COMP_ZSTD])
def test_shim_decompress_reads_the_reference(shim, oracle_ref, method):
    """the reference's cryo_compress (compression.c:125-139) -> cryo_decompress on the GPU."""
    for level in ((1, 50) if method == COMP_LZ4 else (-5, 1, 3)):
        for tag, blk in _blocks():
            comp = oracle_ref.compress(method, level, blk)[0][0]
            rc, out, msg = shim.decompress(method, comp)
            assert rc == 0, (tag, level, rc, msg)
            assert np.array_equal(out, blk), (tag, level)


def test_shim_reads_gucs_at_call_time(shim, oracle_ref):
    """lz4_acceleration_guc / zstd_compression_level_guc are read per call (compression.c:72, :104):
    changing them between two calls changes the stream, and each still round-trips."""
    blk = bg.make_block("M", "lowcard", 9)
    sizes = {}
    for accel in (1, 50):
        shim.set_gucs(COMP_LZ4, accel, 1)
        rc, comp, msg = shim.compress(COMP_LZ4, blk)
        assert rc == 0, msg
        back, ok = oracle_ref.decompress_one(COMP_LZ4, comp)
        assert ok and np.array_equal(back, blk)
        sizes[("lz4", accel)] = len(comp)
    for level in (-5, 3):
        shim.set_gucs(COMP_ZSTD, 1, level)
        rc, comp, msg = shim.compress(COMP_ZSTD, blk)
        assert rc == 0, msg
        back, ok = oracle_ref.decompress_one(COMP_ZSTD, comp)
        assert ok and np.array_equal(back, blk)
        sizes[("zstd", level)] = len(comp)
    assert sizes[("lz4", 1)] < sizes[("lz4", 50)]
    assert sizes[("zstd", 3)] < sizes[("zstd", -5)]
    shim.set_gucs()


def test_shim_malformed_input_returns_false(shim, oracle_ref):
    """cache.c:178-179 maps false to CRYO_ERR_DECOMPRESSION_FAILED; verdicts equal the reference's."""
    blk = bg.make_block("S", "hex", 12)
    for method in (COMP_LZ4, COMP_ZSTD):
        comp = oracle_ref.compress(method, 1, blk)[0][0]
        bad_magic = comp.copy()
        bad_magic[0] ^= 0xFF
        cases = {"truncated": comp[:-100], "trailing": np.concatenate([comp, comp[:3]]),
                 "empty": comp[:0]}
        if method == COMP_ZSTD:
            cases["bad magic"] = bad_magic
        for tag, c in cases.items():
            rc, _, msg = shim.decompress(method, c)
            # the verdict is the reference's own (ZSTD_decompress of an empty input returns 0 bytes and
            # compression.c:116-118 calls that success; LZ4_decompress_safe returns -1)
            want = oracle_ref.decompress_one(method, c)[1]
            assert want is (method == COMP_ZSTD and tag == "empty")
            assert rc == (0 if want else -1), (method, tag, rc, msg)
        # and the context is still good afterwards
        rc, out, _ = shim.decompress(method, comp)
        assert rc == 0 and np.array_equal(out, blk)


def test_shim_unknown_method_raises_elog(shim):
    """compression.c:137 / :157: elog(ERROR, "... unknown compression method")."""
    blk = bg.make_block("S", "hex", 1)
    rc, _, msg = shim.compress(7, blk)
    assert rc == 1 and "unknown compression method" in msg
    rc, _, msg = shim.decompress(7, blk[:100])
    assert rc == 1 and "unknown compression method" in msg
