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
    """The one-warp-per-block decoders (lz4_decode_w.cuh, zstd_decode_w.cuh); the CTA-per-block ones
    are in test_emu_cx.py, the phase-split pipeline further down."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", "libcryoemu.so"))
    for f in (L.emu_lz4w_decode, L.emu_zstdw_decode):
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32)]
        f.restype = C.c_int

    def run(method, stream, cap=MiB, shift=0):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint8)
        sz = C.c_uint32(0)
        fn = L.emu_lz4w_decode if method == 0 else L.emu_zstdw_decode
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


def test_emulated_zstd_encoder_roundtrips_through_libzstd(oracle_ref, oracle_port):
    """zstd_encode.cuh under the emulator: frames must be accepted and restored byte-identically by
    the reference's libzstd (ZSTD_decompress at compression.c:116), sizes within the stated tolerance."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", "libcryoemu.so"))
    L.emu_zstd_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    L.emu_zstd_encode.restype = C.c_int

    def enc(data, level):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = data.size + (data.size >> 8) + 128
        out = np.zeros(cap, dtype=np.uint8)
        sz = C.c_uint32(0)
        st = L.emu_zstd_encode(data.ctypes.data if data.size else None, data.size, out.ctypes.data, cap, level, C.byref(sz))
        assert st == 0, st
        return out[: sz.value].copy()

    for level in (1, -5):
        blk = bg.make_block("S", "hex", 17)
        c = enc(blk, level)
        back, ok = oracle_ref.decompress_one(1, c)
        assert ok and np.array_equal(back, blk)
        ref_size = int(oracle_ref.compress(1, level, blk)[1][0])
        assert len(c) <= 1.15 * ref_size + 64, (level, len(c), ref_size)
    # odd sizes through the plain-C restatement (the reference wrapper only takes 1 MiB blocks)
    src = bg.make_block("D", "lowcard", 2)
    for n in (0, 1, 15, 63, 64, 300, 5000, 65536 + 77):
        c = enc(src[:n], 1)
        got, out = oracle_port.zstd_decode(c, cap=max(n, 1))[:2]
        assert got == n and np.array_equal(out[:n], src[:n]), n


# ---- phase-split pipeline (zstd_decode_p.cuh): stages 1-4 + fallback -------------------------

def _pipeline_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    L = C.CDLL(os.path.join(HERE, "emu", "libcryoemu.so"))
    L.emu_zstdp_decode_multi.restype = C.c_int
    return L


def _run_pipeline(L, comp, cap=MiB, shift=0):
    n = len(comp)
    comp = [np.ascontiguousarray(c, dtype=np.uint8) for c in comp]
    outs = [np.zeros(cap, dtype=np.uint8) for _ in range(n)]
    srcs = (C.c_void_p * n)(*[c.ctypes.data if c.size else None for c in comp])
    dsts = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    csz = (C.c_uint32 * n)(*[c.size for c in comp])
    osz = (C.c_uint32 * n)()
    st = (C.c_int32 * n)(*([-1] * n))
    fl = (C.c_uint32 * n)()
    rc = L.emu_zstdp_decode_multi(n, srcs, csz, dsts, cap, shift, osz, st, fl)
    assert rc == 0, rc
    return list(st), list(osz), outs, list(fl)


@pytest.mark.parametrize("late_prefill", [False, True])
def test_emulated_pipeline_mixed_frames(oracle_ref, late_prefill, monkeypatch):
    """Eleven different frames through the phase-split pipeline in one batch: two entropy groups,
    lanes of one warp on different tables / streams / sequence counts; a truncated frame must
    fail alone (through the fallback), everything libzstd wrote at levels -5..3 must be decoded
    by the pipeline itself (flag 0).  late_prefill: the raw / RLE stage runs after the executor, which
    must then time out waiting for it and write the blocks its matches read itself."""
    if late_prefill:
        monkeypatch.setenv("ZP_EMU_LATE_PREFILL", "1")
    L = _pipeline_lib()
    blocks = [bg.make_block("S", "hex", 31), np.zeros(MiB, dtype=np.uint8), bg.make_block("S", "lowcard", 32),
              bg.make_block("D", "random", 33), bg.regression_block(291, 500), bg.make_block("S", "hex", 34),
              bg.make_block("M", "hex", 35), bg.make_block("S", "random", 36), bg.make_block("S", "hex", 37),
              bg.make_block("M", "lowcard", 38), bg.regression_block(1, 290)]
    # 200 KB of random bytes repeated: a Raw block (written by the raw / RLE stage) that the matches of
    # the Compressed blocks after it copy from, so the executor has to wait for it or write it itself
    rnd = np.random.default_rng(7).integers(0, 256, size=200 * 1024, dtype=np.uint8)
    blocks.append(np.tile(rnd, 6)[:MiB].copy())
    levels = [1, 1, 3, 1, -5, 1, 1, 2, -1, 1, 1, 1]
    comp = [oracle_ref.compress(1, lv, b)[0][0] for lv, b in zip(levels, blocks)]
    comp[5] = comp[5][:-37].copy()
    st, osz, outs, fl = _run_pipeline(L, comp, shift=5)
    for i in range(len(comp)):
        if i == 5:
            assert st[i] != 0 and fl[i] != 0
        else:
            assert st[i] == 0 and osz[i] == MiB and np.array_equal(outs[i], blocks[i]), i
            assert fl[i] == 0, (i, "decoded by the fallback, not by the pipeline")


def test_emulated_pipeline_conformance_and_fallback(oracle_ref):
    """Hand-crafted conformance frames and high-level frames (Repeat_Mode tables, single-stream
    literals): right bytes whichever path takes them; malformed inputs get the warp decoder's verdict."""
    L = _pipeline_lib()
    frames = conformance_frames()
    st, osz, outs, fl = _run_pipeline(L, [f for _, f, _ in frames], cap=4096, shift=1)
    for (name, _, expect), s, z, o in zip(frames, st, osz, outs):
        assert s == 0 and z == len(expect) and bytes(o[: len(expect)]) == expect, name
    blk = bg.make_block("S", "lowcard", 3)
    hi = [oracle_ref.compress(1, lv, blk)[0][0] for lv in (9, 19)]
    st, osz, outs, fl = _run_pipeline(L, hi)
    for s, z, o in zip(st, osz, outs):
        assert s == 0 and z == MiB and np.array_equal(o, blk)
    z = oracle_ref.compress(1, 1, bg.make_block("S", "hex", 5))[0][0]
    bad = z.copy()
    bad[0] ^= 0xFF
    two = np.concatenate([z, z])
    flip = z.copy()
    flip[len(z) // 2] ^= 0x55
    st, osz, outs, fl = _run_pipeline(L, [z[:-100], bad, two, flip, z])
    assert st[0] != 0 and st[1] != 0 and st[2] != 0 and st[4] == 0 and fl[4] == 0
    st2, _, _, _ = _run_pipeline(L, [z], cap=MiB - 16)
    assert st2[0] != 0
    # the flipped byte: whatever the verdict, it must be the warp-per-frame decoder's own
    one = np.zeros(MiB, dtype=np.uint8)
    sz = C.c_uint32(0)
    L.emu_zstdw_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32)]
    ref_st = L.emu_zstdw_decode(flip.ctypes.data, flip.size, one.ctypes.data, MiB, 0, C.byref(sz))
    assert (st[3] == 0) == (ref_st == 0)
    if ref_st == 0:
        assert np.array_equal(outs[3][: sz.value], one[: sz.value])


def test_emulated_pipeline_random_frames(oracle_port):
    """Forty frames of odd sizes and contents (text with repeats, random bytes, runs, mixtures),
    written by libzstd at levels -5..6 into a 96 KiB capacity: several zstd blocks per frame,
    raw / RLE / compressed blocks, treeless literals, all three sequence-table modes.  The
    plain-C restatement of the format (oracle/) is the checker."""
    import benchdata
    L = _pipeline_lib()
    _, zstd = benchdata._libs()
    rng = np.random.default_rng(20260117)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(64)]
    cap = 96 * 1024
    plain, comp = [], []
    for k in range(40):
        n = int(rng.integers(1, cap + 1)) if k % 5 else cap
        kind = k % 4
        if kind == 0:
            buf = b" ".join(words[int(i)] for i in rng.integers(0, 64, size=n // 3 + 1))[:n]
        elif kind == 1:
            buf = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        elif kind == 2:
            buf = b"".join(bytes([int(rng.integers(0, 4))]) * int(rng.integers(1, 3000)) for _ in range(n // 500 + 1))[:n]
        else:
            half = n // 2
            buf = rng.integers(0, 16, size=half, dtype=np.uint8).tobytes() + b"\0" * (n - half)
        buf = buf.ljust(n, b"x")
        src = np.frombuffer(buf, dtype=np.uint8).copy()
        scratch = np.zeros(n + n // 128 + 256, dtype=np.uint8)
        got = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, src.ctypes.data, src.size, int(rng.integers(-5, 7)))
        assert 0 < got <= scratch.size
        plain.append(src)
        comp.append(scratch[:got].copy())
    st, osz, outs, fl = _run_pipeline(L, comp, cap=cap, shift=3)
    for k in range(len(comp)):
        want_n, want = oracle_port.zstd_decode(comp[k], cap=cap)[:2]
        assert want_n == plain[k].size and np.array_equal(want[:want_n], plain[k]), "oracle disagrees with libzstd"
        assert st[k] == 0 and osz[k] == plain[k].size, (k, st[k], osz[k])
        assert np.array_equal(outs[k][: osz[k]], plain[k]), k
    # the pipeline itself should have taken nearly all of them (levels up to 6 rarely use Repeat_Mode)
    assert sum(1 for f in fl if f == 0) >= 30, fl


@pytest.mark.parametrize("late_prefill", [False, True])
def test_emulated_pipeline_decodes_own_encoder_frames(oracle_ref, late_prefill, monkeypatch):
    """Frames written by zstd_encode.cuh (64 KiB zstd blocks, RLE blocks for the zero interior of a
    sparse cryo block): Compressed blocks that are not 128 KiB long in front of RLE blocks, so the
    positions of the raw / RLE stage must come from measured block sizes, not from an assumed
    block length."""
    if late_prefill:
        monkeypatch.setenv("ZP_EMU_LATE_PREFILL", "1")
    L = _pipeline_lib()
    L.emu_zstd_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    L.emu_zstd_encode.restype = C.c_int
    blocks = [bg.make_block("S", "hex", 41), bg.regression_block(1, 290), bg.make_block("M", "hex", 42)]
    comp = []
    for blk in blocks:
        cap = blk.size + (blk.size >> 8) + 128
        out = np.zeros(cap, dtype=np.uint8)
        sz = C.c_uint32(0)
        assert L.emu_zstd_encode(blk.ctypes.data, blk.size, out.ctypes.data, cap, 1, C.byref(sz)) == 0
        comp.append(out[: sz.value].copy())
        back, ok = oracle_ref.decompress_one(1, comp[-1])
        assert ok and np.array_equal(back, blk)
    st, osz, outs, fl = _run_pipeline(L, comp, shift=2)
    for i, blk in enumerate(blocks):
        assert st[i] == 0 and osz[i] == MiB and np.array_equal(outs[i], blk), i
        assert fl[i] == 0, i


@pytest.mark.parametrize("mode", ["jobs", "dropped", "off", "no_server"])
def test_emulated_pipeline_long_runs_handed_to_stage0(oracle_ref, mode, monkeypatch):
    """The ~120 KB zero runs of a sparse frame are jobs stage 4 queues for stage 0 (ZP_JOBS).  'dropped': stage 0
    writes them but never publishes them, as if it had given up waiting -- the check after both kernels must send
    those frames to the warp-per-frame decoder; 'off': stage 4 writes its runs itself; 'no_server': so it does when no CTA
    of stage 0's late pass is running (kernels one after the other, as under a profiler).  Right bytes every time."""
    if mode == "dropped":
        monkeypatch.setenv("ZP_EMU_DROP_JOBS", "1")
    if mode == "off":
        monkeypatch.setenv("ZP_EMU_NO_JOBS", "1")
    if mode == "no_server":
        monkeypatch.setenv("ZP_EMU_NO_SERVER", "1")
    L = _pipeline_lib()
    blocks = [bg.make_block("S", "hex", 41), bg.make_block("S", "lowcard", 42), bg.make_block("M", "hex", 43),
              bg.make_block("S", "hex", 44)]
    comp = [oracle_ref.compress(1, 1, b)[0][0] for b in blocks]
    st, osz, outs, fl = _run_pipeline(L, comp, shift=7)
    for i in range(len(comp)):
        assert st[i] == 0 and osz[i] == MiB and np.array_equal(outs[i], blocks[i]), (mode, i)
    if mode == "dropped":
        assert fl[0] != 0 and fl[3] != 0, fl      # the sparse hex frames have two such runs each
    else:
        assert fl == [0, 0, 0, 0], fl


@pytest.mark.parametrize("late_prefill", [False, True])
def test_emulated_pipeline_matches_around_long_runs(oracle_port, late_prefill, monkeypatch):
    """Blocks built to stress what stage 4 does around the runs it hands to stage 0: long runs of one byte (jobs), of
    several different bytes (the known range holds only the last), data after a run that copies from before it (below
    what stage 0 owes), from inside it, from across its end (a copy that starts in the run and runs out of it), and from
    far behind the ring; raw and RLE blocks in between.  Stage 0's late pass after stage 4 as well, so every wait for it
    times out and stage 4 writes the runs itself."""
    import benchdata
    if late_prefill:
        monkeypatch.setenv("ZP_EMU_LATE_PREFILL", "1")
    L = _pipeline_lib()
    _, zstd = benchdata._libs()
    rng = np.random.default_rng(20260202)
    plain, comp = [], []
    for k in range(12):
        head = rng.integers(0, 256, size=int(rng.integers(300, 3000)), dtype=np.uint8)
        text = rng.integers(97, 103, size=int(rng.integers(2000, 30000)), dtype=np.uint8)
        parts = [head, np.zeros(int(rng.integers(40_000, 300_000)), dtype=np.uint8)]
        parts.append(head[: head.size // 2])                                   # copies from before the first run
        parts.append(np.full(int(rng.integers(33_000, 90_000)), 7 if k % 3 == 0 else 0, dtype=np.uint8))
        parts.append(np.concatenate([np.zeros(40, dtype=np.uint8), text[:200]]))
        parts.append(text)
        parts.append(np.concatenate([np.zeros(24, dtype=np.uint8), text[:64]]))  # across the end of a run, far behind
        if k % 2:
            parts.append(rng.integers(0, 256, size=140_000, dtype=np.uint8))    # a Raw block for stage 0
        parts.append(np.zeros(int(rng.integers(1000, 70_000)), dtype=np.uint8))
        parts.append(text[::-1].copy())
        buf = np.concatenate(parts)
        if buf.size < MiB:
            buf = np.concatenate([buf, np.zeros(MiB - buf.size, dtype=np.uint8)])
        buf = buf[:MiB].copy()
        scratch = np.zeros(MiB + MiB // 128 + 4096, dtype=np.uint8)
        got = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, buf.ctypes.data, buf.size, [1, 3, -1, 2][k % 4])
        assert 0 < got <= scratch.size
        plain.append(buf)
        comp.append(scratch[:got].copy())
    st, osz, outs, fl = _run_pipeline(L, comp, shift=9)
    for k in range(len(comp)):
        want_n, want = oracle_port.zstd_decode(comp[k], cap=MiB)[:2]
        assert want_n == MiB and np.array_equal(want, plain[k]), "oracle disagrees with libzstd"
        assert st[k] == 0 and osz[k] == MiB, (k, st[k], osz[k])
        assert np.array_equal(outs[k], plain[k]), (k, int(np.argmax(outs[k] != plain[k])))


@pytest.mark.parametrize("nframes,cap,seed", [(10, MiB, 1), (11, 384 * 1024, 2), (7, MiB, 3), (13, 200 * 1024, 4)])
def test_emulated_pipeline_random_run_layouts(oracle_port, nframes, cap, seed):
    """Random layouts of long one-byte runs (more of them than a frame may hand to stage 0, of several bytes, some just
    under the job size), random bytes, text, and copies taken from anywhere earlier in the block, at several capacities
    and batch sizes (the emulated grid shares the groups' block indices 1, 3 or 16 ways by the batch size; the early
    pass takes the last half of the batch)."""
    import benchdata
    L = _pipeline_lib()
    _, zstd = benchdata._libs()
    rng = np.random.default_rng(20260300 + seed)
    plain, comp = [], []
    for k in range(nframes):
        parts, size = [], 0
        while size < cap:
            kind = int(rng.integers(0, 6))
            if kind == 0:
                p = np.full(int(rng.integers(20_000, 140_000)), int(rng.integers(0, 3)) * 7, dtype=np.uint8)
            elif kind == 1:
                p = rng.integers(0, 256, size=int(rng.integers(100, 5000)), dtype=np.uint8)
            elif kind == 2:
                p = rng.integers(97, 101, size=int(rng.integers(500, 20000)), dtype=np.uint8)
            elif kind == 3 and size > 100:
                cur = np.concatenate(parts)
                at = int(rng.integers(0, size - 50))
                p = cur[at: at + int(rng.integers(8, 3000))].copy()      # a copy from anywhere earlier
            elif kind == 4:
                p = np.full(int(rng.integers(31_000, 34_000)), 0, dtype=np.uint8)   # around the job size
            else:
                p = rng.integers(0, 256, size=int(rng.integers(60_000, 150_000)), dtype=np.uint8) if rng.integers(0, 4) == 0 \
                    else np.zeros(int(rng.integers(1, 300)), dtype=np.uint8)
            parts.append(p)
            size += p.size
        buf = np.concatenate(parts)[:cap].copy()
        scratch = np.zeros(cap + cap // 128 + 4096, dtype=np.uint8)
        got = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, buf.ctypes.data, buf.size, int(rng.integers(-3, 4)))
        assert 0 < got <= scratch.size
        plain.append(buf)
        comp.append(scratch[:got].copy())
    st, osz, outs, fl = _run_pipeline(L, comp, cap=cap, shift=seed)
    for k in range(nframes):
        want_n, want = oracle_port.zstd_decode(comp[k], cap=cap)[:2]
        assert want_n == cap and np.array_equal(want[:cap], plain[k]), "oracle disagrees with libzstd"
        assert st[k] == 0 and osz[k] == cap, (k, st[k], osz[k])
        assert np.array_equal(outs[k][:cap], plain[k]), (k, int(np.argmax(outs[k][:cap] != plain[k])), fl[k])


def test_emulated_warp_decoders_far_match_wraps_the_ring(emu, oracle_ref):
    """A match short enough for the ring path and so far back that its destination wraps round onto the ring slots
    of its own source (distance + length beyond the 2 KiB ring): the steps of the copy must run in order.  They did
    not (no barrier between them) until the end of round 2; the device hid it by running a warp's lanes in lockstep."""
    rng = np.random.default_rng(20260401)
    blk = np.zeros(MiB, dtype=np.uint8)
    blk[:6000] = rng.integers(97, 101, size=6000, dtype=np.uint8)
    pos = 6000
    for dist, length in [(1500, 200), (1878, 490), (1984, 511), (1985, 300), (2040, 64), (1700, 400), (1990, 480), (1600, 500)]:
        # text of four letters: short matches, so the bytes in front of the far copy are in the ring (a long literal run would
        # have gone to global memory and the copy would have read it from there)
        more = rng.integers(97, 101, size=dist - 200, dtype=np.uint8)
        blk[pos: pos + more.size] = more
        pos += more.size
        blk[pos: pos + length] = blk[pos - dist: pos - dist + length]
        pos += length
    for method, lv in ((0, 1), (1, 1), (1, 3)):
        c = oracle_ref.compress(method, lv, blk)[0][0]
        st, sz, out = emu(method, c, shift=lv)
        assert st == 0 and sz == MiB and np.array_equal(out, blk), (method, lv, int(np.argmax(out != blk)))


@pytest.mark.parametrize("mode, first, last", [("corrupt", 1, 2), ("warp-corrupt", 0, 1), ("cta-corrupt", 2, 3), ("encode", 2, 5)])
def test_emulated_bodies_on_fuzzed_streams(oracle_port, mode, first, last):
    """Batches in which half of the frames carry flipped bytes, a cut or appended bytes (tools/emu_fuzz.py, one seed
    of each mode): every frame's verdict and, when accepted, bytes are the plain-C restatement's; the intact frames of the
    same batch come out right (a rejected frame's jobs and ring never leak into its neighbours).  "encode": the LZ4 and
    zstd encoder bodies over random layouts of 70 001, 4 099 and 131 077 bytes, decoded back by the restatement."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "emu_fuzz.py"), mode, str(first), str(last)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mismatches: 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
