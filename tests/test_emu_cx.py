"""The CTA-per-block decoders on the CPU (not gpu): cryo_cx.cuh + lz4_decode_c.cuh (+ zstd_decode_c.cuh)
under tests/emu/cuda_emu.h, bit-exact against the reference's liblz4 / libzstd and with the same verdicts
on malformed input (LZ4_decompress_safe at compression.c:84).  Two builds: a 64-thread CTA (4 KiB parse
regions: many rounds, chunk cuts and irregular links per block) and the product's 1024 threads."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pg_cryogen_b200 import blockgen as bg

HERE = os.path.dirname(os.path.abspath(__file__))
MiB = 1 << 20


def _lz4():
    L = C.CDLL("liblz4.so.1")
    L.LZ4_compress_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return L


def lz4_compress(b, accel=1):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros(b.size + b.size // 255 + 64, dtype=np.uint8)
    n = _lz4().LZ4_compress_fast(b.ctypes.data, out.ctypes.data, b.size, out.size, accel)
    assert n > 0
    return out[:n].copy()


def lz4_reference(c, cap):
    """LZ4_decompress_safe as compression.c:84 calls it: (bytes or negative, output)."""
    c = np.ascontiguousarray(c, dtype=np.uint8)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    n = _lz4().LZ4_decompress_safe(c.ctypes.data, out.ctypes.data, c.size, cap)
    return n, out[:cap]


@pytest.fixture(scope="module", params=["64", "1024"])
def cx(request):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", f"libcryoemu_cx{request.param}.so"))
    L.emu_lz4c_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32)]
    assert L.emu_cx_threads() == int(request.param)

    def run(stream, cap, shift=0):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(max(cap, 16), dtype=np.uint8)
        sz = C.c_uint32(0)
        st = L.emu_lz4c_decode(s.ctypes.data if s.size else None, s.size, out.ctypes.data, cap, shift, C.byref(sz))
        return st, sz.value, out[:cap]
    run.threads = int(request.param)
    return run


def _slices(n):
    d, s = bg.make_block("D", "lowcard", 2), bg.make_block("S", "hex", 3)
    m, h = bg.make_block("M", "hex", 4), bg.make_block("D", "hex", 5)
    r = np.frombuffer(bg.rand_bytes(7, n).tobytes(), dtype=np.uint8)
    return [("D/lowcard", d[-n:]), ("S/hex", np.concatenate([s[: n // 2], s[-(n - n // 2):]])), ("M/hex", m[-n:]),
            ("D/hex", h[-n:]), ("random", r), ("zeros", np.zeros(n, dtype=np.uint8))]


@pytest.mark.parametrize("n", [13, 100, 4097, 70000])
def test_cx_lz4_matches_liblz4(cx, n):
    if cx.threads == 1024 and n > 5000:
        n = 20000
    for tag, blk in _slices(n):
        blk = np.ascontiguousarray(blk)
        for accel in (1, 50):
            c = lz4_compress(blk, accel)
            st, sz, out = cx(c, n, shift=(n + accel) % 16)
            assert st == 0 and sz == n and np.array_equal(out, blk), (tag, n, accel, st, sz)


def test_cx_lz4_full_blocks(cx, oracle_ref):
    kinds = [("S", "hex"), ("M", "lowcard")] if cx.threads == 1024 else [("S", "hex"), ("M", "lowcard"), ("D", "lowcard"), ("D", "hex")]
    for kind, pl in kinds:
        blk = bg.make_block(kind, pl, 31)
        c = oracle_ref.compress(0, 1, blk)[0][0]
        st, sz, out = cx(c, MiB, shift=5)
        assert st == 0 and sz == MiB and np.array_equal(out, blk), (kind, pl, st, sz)


def test_cx_lz4_malformed_verdicts_equal_the_reference(cx):
    blk = np.ascontiguousarray(bg.make_block("D", "lowcard", 9)[-30000:])
    c = lz4_compress(blk)
    far = np.array([0x10, 65, 5, 0, 0x50, 97, 98, 99, 100, 101], dtype=np.uint8)
    cases = {"valid": (c, 30000), "truncated-100": (c[:-100], 30000), "truncated-1": (c[:-1], 30000),
             "trailing": (np.concatenate([c, np.array([1, 2, 3], dtype=np.uint8)]), 30000),
             "offset-before-start": (far, 30000), "one-byte": (c[:1], 30000), "empty": (c[:0], 30000),
             "cap-1": (c, 29999), "cap-16": (c, 30000 - 16), "cap+16": (c, 30016),
             "short-output": (np.array([0x50, 1, 2, 3, 4, 5], dtype=np.uint8), 30000)}
    for tag, (s, cap) in cases.items():
        want, ref_out = lz4_reference(s, cap)
        st, sz, out = cx(s, cap)
        assert (st == 0) == (want >= 0), (tag, st, want)
        if want >= 0:
            assert sz == want and np.array_equal(out[:want], ref_out[:want]), tag


def test_cx_lz4_mutated_streams(cx):
    """Single-byte corruptions: the verdict is liblz4's; where it accepts, so do we, with the same bytes.
    (Known deviation, DESIGN.md section 1: a match offset of 0 is rejected here, liblz4 1.9.4 accepts it.)"""
    rng = np.random.default_rng(5)
    blk = np.ascontiguousarray(bg.make_block("D", "lowcard", 3)[-6000:])
    c = lz4_compress(blk)
    rounds = 24 if cx.threads == 1024 else 120
    for k in range(rounds):
        m = c.copy()
        pos = int(rng.integers(0, m.size))
        m[pos] = int(rng.integers(0, 256))
        want, ref_out = lz4_reference(m, 6000)
        st, sz, out = cx(m, 6000, shift=k % 16)
        if want >= 0 and st == 3:
            continue                    # offset 0
        assert (st == 0) == (want >= 0), (k, pos, st, want)
        if want >= 0:
            assert sz == want and np.array_equal(out[:want], ref_out[:want]), (k, pos)
