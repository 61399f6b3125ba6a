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


# ---------------------------------------------------------------- zstd: CTA-per-frame stage 4 (zstd_decode_c.cuh)

def zstd_compress(b, level=1):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.zeros(b.size + b.size // 128 + 512, dtype=np.uint8)
    n = z.ZSTD_compress(out.ctypes.data, out.size, b.ctypes.data, b.size, level)
    assert n < out.size
    return out[:n].copy()


@pytest.fixture(scope="module", params=["64", "1024"])
def zcx(request):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", f"libcryoemu_cx{request.param}.so"))
    L.emu_zstdc_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint32)]

    def run(stream, cap, shift=0):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(max(cap, 16), dtype=np.uint8)
        sz, flag = C.c_uint32(0), C.c_uint32(0)
        st = L.emu_zstdc_decode(s.ctypes.data if s.size else None, s.size, out.ctypes.data, cap, shift, C.byref(sz), C.byref(flag))
        return st, sz.value, out[:cap], flag.value
    run.threads = int(request.param)
    return run


def test_cx_zstd_matches_libzstd(zcx, oracle_port):
    """Frames written by the reference's libzstd (compression.c:93-109), decoded by stages 1-3 + the CTA stage 4."""
    sizes = (100, 4097, 70000, 300000) if zcx.threads == 64 else (4097, 140000)
    for n in sizes:
        for tag, blk in _slices(n):
            blk = np.ascontiguousarray(blk)
            for level in (1, 3, -5):
                c = zstd_compress(blk, level)
                st, sz, out, flag = zcx(c, n, shift=(n + level) % 16)
                assert st == 0 and sz == n and np.array_equal(out, blk), (tag, n, level, st, sz, flag)
                # the pipeline's work area holds cap / 8 + 64 sequence records per frame (zp_seq_cap); a frame with
                # more is decoded by the warp-per-frame decoder (flag), any other must take the CTA stage
                nseq = oracle_port.zstd_decode(c, n, stats=True)[2]["sequences"]
                assert flag == (1 if nseq > n // 8 + 64 else 0), (tag, n, level, nseq, "routing")


def test_cx_zstd_full_blocks(zcx, oracle_ref):
    kinds = [("S", "hex")] if zcx.threads == 1024 else [("S", "hex"), ("M", "lowcard"), ("D", "lowcard")]
    for kind, pl in kinds:
        blk = bg.make_block(kind, pl, 31)
        c = oracle_ref.compress(1, 1, blk)[0][0]
        st, sz, out, flag = zcx(c, MiB, shift=3)
        assert st == 0 and sz == MiB and np.array_equal(out, blk) and flag == 0, (kind, pl, st, sz, flag)


def test_cx_zstd_mutated_frames(zcx, oracle_ref, oracle_port):
    """Corrupted frames: never a crash or a write outside the block; the verdict is the port's
    (DESIGN.md section 1 lists where the port is stricter than libzstd), accepted frames give its bytes."""
    rng = np.random.default_rng(11)
    blk = np.ascontiguousarray(bg.make_block("D", "lowcard", 3)[-9000:])
    c = zstd_compress(blk, 1)
    for k in range(20 if zcx.threads == 1024 else 80):
        m = c.copy()
        pos = int(rng.integers(0, m.size))
        m[pos] = int(rng.integers(0, 256))
        want_n, want = oracle_port.zstd_decode(m, 9000)
        st, sz, out, flag = zcx(m, 9000, shift=k % 16)
        assert st != -100
        assert (st == 0) == (want_n >= 0), (k, pos, st, want_n)
        if want_n >= 0:
            assert sz == want_n and np.array_equal(out[:sz], want[:sz]), (k, pos)
