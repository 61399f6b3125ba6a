"""Kernel bodies on the CPU (not gpu): the .cuh decoders run under tests/emu/cuda_emu.h
(fibers + cooperative barriers) and must be bit-exact against the oracle.  This checks
the barrier protocol, warp intrinsics use, shared-memory indexing and bounds before any
GPU time is spent.  It is test scaffolding, not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pg_cryogen_b200 import blockgen as bg

import golden_util
from zstd_vectors import conformance_frames

HERE = os.path.dirname(os.path.abspath(__file__))
MiB = 1 << 20


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", "libcryoemu.so"))
    for f in (L.emu_lz4_decode, L.emu_zstd_decode):
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32)]
        f.restype = C.c_int

    def run(method, stream, cap=MiB, shift=0):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint8)
        sz = C.c_uint32(0)
        fn = L.emu_lz4_decode if method == 0 else L.emu_zstd_decode
        st = fn(s.ctypes.data if s.size else None, s.size, out.ctypes.data, cap, shift, C.byref(sz))
        return st, sz.value, out
    return run


def test_emulated_decoders_on_golden_streams(emu):
    import hashlib
    for name, method, level, stream, sha in golden_util.load():
        if name.startswith(("M_", "D_")):
            continue                      # match-heavy 1 MiB blocks are slow under emulation
        st, sz, out = emu(method, stream, shift=len(name) % 16)
        assert st == 0 and sz == MiB, (name, st)
        assert hashlib.sha256(out.tobytes()).digest() == sha, name


@pytest.mark.parametrize("kind,payload", [("S", "hex"), ("S", "lowcard"), ("M", "hex"), ("D", "random")])
def test_emulated_decoders_match_reference(emu, oracle_ref, kind, payload):
    blk = bg.make_block(kind, payload, 21)
    for method, levels in ((0, (1, 50)), (1, (-5, 1, 3))):
        for lv in levels:
            c = oracle_ref.compress(method, lv, blk)[0][0]
            st, sz, out = emu(method, c, shift=lv % 16)
            assert st == 0 and sz == MiB and np.array_equal(out, blk), (kind, payload, method, lv)


def test_emulated_zstd_conformance_vectors(emu):
    for name, frame, expect in conformance_frames():
        st, sz, out = emu(1, frame, shift=1)
        assert st == 0 and sz == len(expect) and bytes(out[: len(expect)]) == expect, name


def test_emulated_decoders_reject_malformed(emu, oracle_ref):
    blk = bg.make_block("S", "hex", 5)
    c = oracle_ref.compress(0, 1, blk)[0][0]
    assert emu(0, c[:-100])[0] != 0
    assert emu(0, np.concatenate([c, c[:3]]))[0] != 0
    assert emu(0, c, cap=MiB - 16)[0] != 0
    z = oracle_ref.compress(1, 1, blk)[0][0]
    assert emu(1, z[:-100])[0] != 0
    assert emu(1, z, cap=MiB - 16)[0] != 0
    bad = z.copy()
    bad[0] ^= 0xFF
    assert emu(1, bad)[0] != 0
