"""-m gpu: batched zstd frame decompression (k_zstd_decode) against the oracle.

Replaces ZSTD_decompress at reference compression.c:116.  Bit-exact on every frame
the reference's compressor can emit (levels -5..22, compression.c:53) for all block
kinds and payloads, plus hand-crafted conformance frames for format branches libzstd
never emits and malformed inputs.
"""
import os

import numpy as np
import pytest

from pg_cryogen_b200 import COMP_LZ4, COMP_ZSTD, CRYO_BLCKSZ
from pg_cryogen_b200 import blockgen as bg

from gpu_util import decode_device
from zstd_vectors import conformance_frames

pytestmark = pytest.mark.gpu


def _blocks():
    blocks, tags = [], []
    for kind in "SMD":
        for pl in bg.PAYLOADS:
            blocks.append(bg.make_block(kind, pl, 11))
            tags.append(f"{kind}/{pl}")
    blocks.append(np.zeros(CRYO_BLCKSZ, dtype=np.uint8))
    tags.append("zeros")
    blocks.append(bg.regression_block(1, 290))
    tags.append("regression-1")
    blocks.append(bg.regression_block(291, 500))
    tags.append("regression-2")
    # 200 KB of random bytes repeated: Raw blocks whose bytes the matches of later blocks copy (the
    # pipeline's executor then depends on its raw / RLE stage)
    rnd = np.random.default_rng(7).integers(0, 256, size=200 * 1024, dtype=np.uint8)
    blocks.append(np.tile(rnd, 6)[:CRYO_BLCKSZ].copy())
    tags.append("random-repeated")
    return np.stack(blocks), tags


@pytest.mark.parametrize("levels", [(-5, -3, -1), (1, 2, 3), (4, 6, 9), (12, 19, 22)])
def test_zstd_decode_bit_exact(gpu, oracle_ref, levels):
    blocks, tags = _blocks()
    chunks, want, names = [], [], []
    for lv in levels:
        comp, _, _ = oracle_ref.compress(COMP_ZSTD, lv, blocks, nthreads=8)
        for i, c in enumerate(comp):
            chunks.append(c)
            want.append(i)
            names.append(f"{tags[i]}@{lv}")
    out, osz, st = decode_device(gpu, COMP_ZSTD, chunks)
    for k in range(len(chunks)):
        assert st[k] == 0, (names[k], st[k])
        assert osz[k] == CRYO_BLCKSZ, names[k]
        assert np.array_equal(out[k], blocks[want[k]]), names[k]


def test_zstd_decode_conformance_vectors(gpu, oracle_ref):
    """Format branches ZSTD_compress never emits; each vector is first confirmed with
    the reference's own decompressor, then must decode identically on the GPU."""
    vecs = conformance_frames()
    out, osz, st = decode_device(gpu, COMP_ZSTD, [v for _, v, _ in vecs])
    for k, (name, frame, expect) in enumerate(vecs):
        ref_out, ref_ok = oracle_ref.decompress_one(COMP_ZSTD, frame)
        assert ref_ok, name
        assert st[k] == 0, (name, st[k])
        assert osz[k] == len(expect), name
        assert bytes(out[k, : len(expect)]) == expect == bytes(ref_out[: len(expect)]), name


def test_zstd_decode_mixed_methods_in_one_batch(gpu, oracle_ref):
    """storage.h:64: the method is per block; sql/pg_cryogen.sql:26-28 mixes them."""
    blocks = np.stack([bg.regression_block(1, 290), bg.regression_block(291, 500),
                       bg.regression_block(501, 790), bg.regression_block(791, 1000)])
    z, _, _ = oracle_ref.compress(COMP_ZSTD, 1, blocks[:2])
    l, _, _ = oracle_ref.compress(COMP_LZ4, 1, blocks[2:])
    out, osz, st = decode_device(gpu, [COMP_ZSTD, COMP_ZSTD, COMP_LZ4, COMP_LZ4], z + l)
    assert (st == 0).all()
    assert np.array_equal(out, blocks)


def test_zstd_decode_malformed_matches_reference_verdict(gpu, oracle_ref):
    blk = bg.make_block("S", "hex", 5)
    c = oracle_ref.compress(COMP_ZSTD, 1, blk)[0][0]
    bad_magic = c.copy()
    bad_magic[0] ^= 0xFF
    reserved_block = c.copy()
    reserved_block[9] |= 0x06            # first block header (10-byte frame header): type 3
    wrong_fcs = c.copy()
    wrong_fcs[6] ^= 0x01                 # FCS no longer 1 MiB
    flipped = c.copy()
    flipped[len(c) // 2] ^= 0x55
    half = oracle_ref.compress(COMP_ZSTD, 1, blk)[0][0]
    cases = [
        ("valid", c),
        ("truncated-100", c[:-100]),
        ("truncated-1", c[:-1]),
        ("trailing-garbage", np.concatenate([c, np.array([1, 2, 3], dtype=np.uint8)])),
        ("bad-magic", bad_magic),
        ("reserved-block-type", reserved_block),
        ("wrong-content-size", wrong_fcs),
        ("empty", np.zeros(0, dtype=np.uint8)),
        ("header-only", c[:8]),
    ]
    out, osz, st = decode_device(gpu, COMP_ZSTD, [x for _, x in cases])
    for k, (name, x) in enumerate(cases):
        _, ref_ok = oracle_ref.decompress_one(COMP_ZSTD, x)
        assert (st[k] == 0) == ref_ok, (name, st[k], ref_ok)
    assert np.array_equal(out[0], blk)
    # a bit flip in the middle: the reference may or may not notice; we only require
    # that the GPU never reports success with different bytes than the reference
    o2, _, s2 = decode_device(gpu, COMP_ZSTD, [flipped])
    r2, ok2 = oracle_ref.decompress_one(COMP_ZSTD, flipped)
    if s2[0] == 0 and ok2:
        assert np.array_equal(o2[0], r2)


def test_zstd_decode_capacity_too_small(gpu, oracle_ref):
    blk = bg.make_block("M", "hex", 2)
    c = oracle_ref.compress(COMP_ZSTD, 1, blk)[0][0]
    out, osz, st = decode_device(gpu, COMP_ZSTD, [c], block_size=CRYO_BLCKSZ - 16)
    assert st[0] != 0


def test_zstd_decode_concatenated_and_skippable_frames(gpu, oracle_ref):
    """ZSTD_decompress accepts these; compression.c:102 never writes them."""
    import ctypes as C
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    blk = bg.make_block("M", "lowcard", 9)
    halves = []
    for h in (blk[: CRYO_BLCKSZ // 2], blk[CRYO_BLCKSZ // 2:]):
        h = np.ascontiguousarray(h)
        dst = np.zeros(h.size + 4096, dtype=np.uint8)
        n = z.ZSTD_compress(dst.ctypes.data, dst.size, h.ctypes.data, h.size, 1)
        halves.append(dst[:n].copy())
    skippable = np.frombuffer(b"\x50\x2a\x4d\x18\x05\x00\x00\x00hello", dtype=np.uint8)
    two = np.concatenate(halves)
    skip = np.concatenate([skippable, halves[0], skippable, halves[1]])
    out, osz, st = decode_device(gpu, COMP_ZSTD, [two, skip])
    for k, frame in enumerate((two, skip)):
        r, ok = oracle_ref.decompress_one(COMP_ZSTD, frame)
        assert ok and st[k] == 0 and osz[k] == CRYO_BLCKSZ
        assert np.array_equal(out[k], blk) and np.array_equal(r, blk)


def test_zstd_decode_host_api(gpu, oracle_ref):
    blocks = np.stack([bg.make_block(k, "lowcard", 30 + i) for i, k in enumerate("SMDS")])
    comp, _, _ = oracle_ref.compress(COMP_ZSTD, 1, blocks)
    out, osz, st = gpu.decompress_host(COMP_ZSTD, comp)
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    assert np.array_equal(out, blocks)


def _host_batch(oracle_ref):
    blocks = [bg.make_block("S", "hex", 60), bg.make_block("S", "lowcard", 61), bg.make_block("M", "hex", 62),
              bg.make_block("D", "hex", 63), np.zeros(CRYO_BLCKSZ, dtype=np.uint8),
              bg.make_block("D", "random", 64), bg.regression_block(1, 290), bg.make_block("M", "lowcard", 65)]
    blocks = np.stack(blocks * 3)                       # 24 blocks
    methods = np.array([i % 2 for i in range(len(blocks))], dtype=np.int32)
    comp = [oracle_ref.compress(int(m), 1, b)[0][0] for m, b in zip(methods, blocks)]
    return blocks, methods, comp


def test_host_api_sparse_return_is_bit_exact(gpu, oracle_ref):
    """cryogpu_decompress_host ships only the non-zero 4 KiB pages and zero-fills the rest on the
    host; the caller's blocks must come out bit-exact whether they are sparse, dense or all zero,
    and a pre-filled destination must be fully overwritten."""
    blocks, methods, comp = _host_batch(oracle_ref)
    out = np.full(blocks.shape, 0xA5, dtype=np.uint8)
    got, osz, st = gpu.decompress_host(methods, comp, out=out)
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    assert np.array_equal(got, blocks)
    h2d, d2h = gpu.last_transfer_bytes()
    assert h2d >= sum(len(c) for c in comp)
    assert d2h < blocks.size                            # the sparse blocks did not cross the bus whole


def test_host_api_dense_return_env_switch(oracle_ref, monkeypatch):
    from pg_cryogen_b200 import CryoGPU
    monkeypatch.setenv("CRYOGPU_SPARSE_D2H", "0")
    g = CryoGPU(0)
    try:
        blocks, methods, comp = _host_batch(oracle_ref)
        got, osz, st = g.decompress_host(methods, comp)
        assert (st == 0).all() and np.array_equal(got, blocks)
        assert g.last_transfer_bytes()[1] >= blocks.size
    finally:
        g.close()


def test_host_api_dense_batch_takes_the_plain_copy(gpu, oracle_ref):
    blocks = np.stack([bg.make_block("D", "hex", 80 + i) for i in range(6)])
    comp = [oracle_ref.compress(1, 1, b)[0][0] for b in blocks]
    got, osz, st = gpu.decompress_host([1] * 6, comp)
    assert (st == 0).all() and np.array_equal(got, blocks)
    assert gpu.last_transfer_bytes()[1] >= blocks.size


def test_zstd_pipeline_takes_what_libzstd_writes(gpu, oracle_ref):
    """The phase-split pipeline (zstd_decode_p.cuh) must itself decode every frame libzstd writes
    for cryo blocks at levels -5..3 (no hand-over to the one-warp-per-frame decoder), and hand
    over exactly the frames it cannot take: Repeat_Mode tables (level 19 here), a truncated frame,
    two concatenated frames.  Whatever path decodes a frame, bytes and status are the same."""
    blocks, tags = _blocks()
    chunks, want = [], []
    for lv in (-5, -1, 1, 2, 3):
        comp, _, _ = oracle_ref.compress(COMP_ZSTD, lv, blocks, nthreads=8)
        chunks += comp
        want += list(range(len(comp)))
    out, osz, st = decode_device(gpu, COMP_ZSTD, chunks)
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    for k in range(len(chunks)):
        assert np.array_equal(out[k], blocks[want[k]]), k
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == len(chunks) and fallback == 0, (frames, fallback)

    z = oracle_ref.compress(COMP_ZSTD, 1, blocks[:1])[0][0]
    hi = oracle_ref.compress(COMP_ZSTD, 19, bg.make_block("S", "lowcard", 3)[None, :])[0][0]
    odd = [z, z[:-100].copy(), np.concatenate([z, z]), hi, z]
    out, osz, st = decode_device(gpu, [COMP_ZSTD, COMP_ZSTD, COMP_ZSTD, COMP_ZSTD, COMP_LZ4], odd)
    assert st[0] == 0 and st[1] != 0 and st[2] != 0 and st[3] == 0 and st[4] != 0
    assert np.array_equal(out[0], blocks[0]) and np.array_equal(out[3], bg.make_block("S", "lowcard", 3))
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == 4 and 2 <= fallback <= 3, (frames, fallback)


def test_zstd_pipeline_large_batch_mixed_kinds(gpu, oracle_ref):
    """A batch wide enough for several entropy groups per SM and for the raw / RLE stage to run
    beside the executor: 600 frames cycling through every block kind and two levels."""
    blocks, tags = _blocks()
    comp1, _, _ = oracle_ref.compress(COMP_ZSTD, 1, blocks, nthreads=8)
    comp3, _, _ = oracle_ref.compress(COMP_ZSTD, -3, blocks, nthreads=8)
    pool = [(c, i) for i, c in enumerate(comp1)] + [(c, i) for i, c in enumerate(comp3)]
    pick = [pool[(7 * k + k // 5) % len(pool)] for k in range(600)]
    out, osz, st = decode_device(gpu, COMP_ZSTD, [c for c, _ in pick])
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    for k, (_, i) in enumerate(pick):
        assert np.array_equal(out[k], blocks[i]), (k, tags[i])
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == 600 and fallback == 0


def test_zstd_pipeline_random_frames_small_capacity(gpu, oracle_port):
    """200 frames of odd sizes and contents at levels -5..6 in a 96 KiB capacity (block_size is a
    run-time parameter of the library): several zstd blocks per frame, raw / RLE / compressed
    blocks, treeless literals, all sequence-table modes; the plain-C restatement is the checker."""
    import benchdata
    _, zstd = benchdata._libs()
    rng = np.random.default_rng(20260118)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(64)]
    cap = 96 * 1024
    plain, comp = [], []
    for k in range(200):
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
        src = np.frombuffer(buf.ljust(n, b"x"), dtype=np.uint8).copy()
        scratch = np.zeros(n + n // 128 + 256, dtype=np.uint8)
        got = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, src.ctypes.data, src.size, int(rng.integers(-5, 7)))
        assert 0 < got <= scratch.size
        plain.append(src)
        comp.append(scratch[:got].copy())
    out, osz, st = decode_device(gpu, COMP_ZSTD, comp, block_size=cap)
    for k in range(len(comp)):
        assert st[k] == 0 and osz[k] == plain[k].size, (k, st[k], osz[k])
        assert np.array_equal(out[k][: osz[k]], plain[k]), k
    for k in (0, 1, 2, 3, 7):
        want_n, want = oracle_port.zstd_decode(comp[k], cap=cap)[:2]
        assert want_n == plain[k].size and np.array_equal(want[:want_n], plain[k])
    frames, fallback = gpu.zstd_pipeline_stats()
    assert frames == len(comp) and fallback <= len(comp) // 4, (frames, fallback)


def test_host_api_zero_runs_returned_to_the_kernel(gpu, oracle_ref):
    """cryogpu_set_zero_by_unmap: for destinations in private anonymous memory the zero runs of sparse blocks
    are given back to the kernel (madvise) instead of being written; the caller's blocks must read exactly as
    the reference's output whatever they held before, with the option on and off, and for dense blocks too."""
    blocks = np.stack([bg.make_block("S", "hex", 90 + i) for i in range(10)] +
                      [bg.make_block("M", "lowcard", 77), bg.make_block("D", "hex", 78), np.zeros(CRYO_BLCKSZ, dtype=np.uint8)])
    methods = [i & 1 for i in range(blocks.shape[0])]
    comp = [oracle_ref.compress(m, 1, b)[0][0] for m, b in zip(methods, blocks)]
    for on in (1, 0, 1):
        assert gpu.lib.cryogpu_set_zero_by_unmap(gpu.handle, on) == 0
        out = np.full((blocks.shape[0], CRYO_BLCKSZ), 0x5A, dtype=np.uint8)     # pageable, full of stale bytes
        out, osz, st = gpu.decompress_host(methods, comp, out=out)
        assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
        assert np.array_equal(out, blocks), on
        # and again into the same (now partly unmapped) destination
        out[:, ::4096] = 0xA5
        out, osz, st = gpu.decompress_host(methods, comp, out=out)
        assert (st == 0).all() and np.array_equal(out, blocks), on
    gpu.lib.cryogpu_set_zero_by_unmap(gpu.handle, 0)


@pytest.mark.parametrize("knobs", [{}, {"CRYOGPU_ZP_JOBS": "0", "CRYOGPU_ZP_EARLY_CTAS": "0"},
                                   {"CRYOGPU_ZP_JOBS": "0"}, {"CRYOGPU_ZP_EARLY_CTAS": "0"},
                                   {"CRYOGPU_ZP_JOBS": "2", "CRYOGPU_ZP_EARLY_PCT": "100", "CRYOGPU_ZP_PF_INFLIGHT": "1"}])
def test_zstd_pipeline_jobs_and_early_pass_settings(knobs):
    """The executor's long runs as jobs of the raw / RLE stage and that stage's early pass at guessed positions, in every
    combination of their switches (read once per process, hence a fresh one each: tests/zp_knobs_case.py): 400 frames,
    a quarter of them with every guess wrong, bit-exact and none left to the fallback decoder."""
    import subprocess
    import sys
    env = dict(os.environ)
    env.update(knobs)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "zp_knobs_case.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
