"""-m gpu: batched LZ4 block decompression (k_lz4_decode) against the oracle.

Replaces LZ4_decompress_safe at reference compression.c:84.  Bit-exact on valid
streams; on malformed streams the per-block status must be non-zero exactly where
the reference's cryo_decompress returns false, and nothing outside
the block's output slot may be written.
"""
import numpy as np
import pytest

from pg_cryogen_b200 import COMP_LZ4, CRYO_BLCKSZ
from pg_cryogen_b200 import blockgen as bg

from gpu_util import decode_device

pytestmark = pytest.mark.gpu


def _cases():
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
    return np.stack(blocks), tags


def test_lz4_decode_bit_exact_all_kinds_and_accelerations(gpu, oracle_ref):
    blocks, tags = _cases()
    chunks, want, names = [], [], []
    for accel in (0, 1, 2, 5, 10, 25, 50):
        comp, _, _ = oracle_ref.compress(COMP_LZ4, accel, blocks)
        for i, c in enumerate(comp):
            chunks.append(c)
            want.append(i)
            names.append(f"{tags[i]}@{accel}")
    out, osz, st = decode_device(gpu, COMP_LZ4, chunks)
    for k in range(len(chunks)):
        assert st[k] == 0, (names[k], st[k])
        assert osz[k] == CRYO_BLCKSZ, names[k]
        assert np.array_equal(out[k], blocks[want[k]]), names[k]


def test_lz4_decode_unaligned_sources(gpu, oracle_ref):
    """Compressed blocks at every byte alignment inside the source buffer."""
    import torch
    blk = bg.make_block("M", "lowcard", 3)
    comp = oracle_ref.compress(COMP_LZ4, 1, blk)[0][0]
    n = 16
    stride = (len(comp) + 64) & ~15
    buf = np.zeros(n * stride + 64, dtype=np.uint8)
    offs = np.array([i * stride + i for i in range(n)], dtype=np.uint64)
    for i in range(n):
        buf[int(offs[i]): int(offs[i]) + len(comp)] = comp
    dev = torch.device("cuda", gpu.device)
    d_src = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_sz = torch.full((n,), len(comp), dtype=torch.int32, device=dev)
    d_me = torch.zeros((n,), dtype=torch.int32, device=dev)
    d_dst = torch.zeros((n, CRYO_BLCKSZ), dtype=torch.uint8, device=dev)
    d_osz = torch.zeros((n,), dtype=torch.int32, device=dev)
    d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    gpu.decompress_device(d_me, d_src, d_off, d_sz, d_dst, CRYO_BLCKSZ, d_osz, d_st, n,
                          stream=torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    assert (d_st.cpu().numpy() == 0).all()
    out = d_dst.cpu().numpy()
    for i in range(n):
        assert np.array_equal(out[i], blk), i


def _malformed(comp):
    """(name, stream) pairs derived from a valid stream, as in SURVEY.md D.1."""
    c = np.asarray(comp, dtype=np.uint8)
    # 1 literal, then a match at offset 5 while only 1 byte has been produced
    far = np.array([0x10, 65, 5, 0, 0x50, 97, 98, 99, 100, 101], dtype=np.uint8)
    return [
        ("valid", c),
        ("truncated-100", c[:-100]),
        ("truncated-1", c[:-1]),
        ("trailing-garbage", np.concatenate([c, np.array([1, 2, 3], dtype=np.uint8)])),
        ("offset-before-start", far),
        ("one-byte", c[:1]),
    ]


def test_lz4_decode_malformed_matches_reference_verdict(gpu, oracle_ref):
    blk = bg.make_block("S", "hex", 5)
    comp = oracle_ref.compress(COMP_LZ4, 1, blk)[0][0]
    cases = _malformed(comp)
    chunks = [c for _, c in cases]
    out, osz, st = decode_device(gpu, COMP_LZ4, chunks, fill=0x5A)
    for k, (name, c) in enumerate(cases):
        _, ref_ok = oracle_ref.decompress_one(COMP_LZ4, c)
        assert (st[k] == 0) == ref_ok, (name, st[k], ref_ok)
    assert np.array_equal(out[0], blk)


def test_lz4_decode_empty_input_fails(gpu):
    out, osz, st = decode_device(gpu, COMP_LZ4, [np.zeros(0, dtype=np.uint8)])
    assert st[0] != 0        # LZ4_decompress_safe returns -1 on empty input


def test_lz4_decode_capacity_too_small(gpu, oracle_ref):
    """Output capacity one byte short: the reference fails."""
    blk = bg.make_block("M", "hex", 2)
    comp = oracle_ref.compress(COMP_LZ4, 1, blk)[0][0]
    out, osz, st = decode_device(gpu, COMP_LZ4, [comp], block_size=CRYO_BLCKSZ - 16)
    assert st[0] != 0


def test_lz4_decode_short_output_is_accepted(gpu, oracle_port):
    """A valid stream that decodes to 5 bytes: LZ4_decompress_safe returns 5 and the
    reference accepts it (compression.c:85-88, Assert compiled out)."""
    stream = np.array([0x50, 1, 2, 3, 4, 5], dtype=np.uint8)
    out, osz, st = decode_device(gpu, COMP_LZ4, [stream])
    assert st[0] == 0 and osz[0] == 5
    assert out[0, :5].tolist() == [1, 2, 3, 4, 5]


def test_lz4_decode_host_api_roundtrip(gpu, oracle_ref):
    blocks = np.stack([bg.make_block(k, "hex", 20 + i) for i, k in enumerate("SMDS")])
    comp, _, _ = oracle_ref.compress(COMP_LZ4, 1, blocks)
    out, osz, st = gpu.decompress_host(COMP_LZ4, comp)
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    assert np.array_equal(out, blocks)


def test_unknown_method_is_reported_per_block(gpu, oracle_ref):
    blk = bg.make_block("S", "hex", 1)
    comp = oracle_ref.compress(COMP_LZ4, 1, blk)[0][0]
    out, osz, st = decode_device(gpu, [COMP_LZ4, 7], [comp, comp])
    assert st[0] == 0 and st[1] == 6


def test_lz4_decode_routed_batch_uses_both_decoders(gpu, oracle_ref):
    """More than two blocks per SM: every block is routed by its estimated sequence count.  Sparse and
    literal-heavy blocks stay with the warp-per-block kernel, dense low-cardinality blocks (120 K
    sequences each) go to the CTA-per-block kernel; both must be bit-exact."""
    kinds = [("S", "hex"), ("D", "lowcard"), ("M", "hex"), ("D", "hex"), ("S", "lowcard"), ("M", "lowcard")]
    uniq = np.stack([bg.make_block(k, p, 50 + i) for i, (k, p) in enumerate(kinds)])
    comp = oracle_ref.compress(COMP_LZ4, 1, uniq, nthreads=6)[0]
    n = 2 * 148 + 37
    chunks = [comp[i % len(comp)] for i in range(n)]
    out, osz, st = decode_device(gpu, COMP_LZ4, chunks)
    assert (st == 0).all() and (osz == CRYO_BLCKSZ).all()
    for i in range(n):
        assert np.array_equal(out[i], uniq[i % len(comp)]), i
    total, cta = gpu.lz4_route_stats()
    assert total == n
    per = {kinds[j]: sum(1 for i in range(n) if i % len(comp) == j) for j in range(len(kinds))}
    assert cta >= per[("D", "lowcard")], (total, cta)          # the match-rich kind is on the CTA kernel
    assert cta <= n - per[("S", "hex")] - per[("D", "hex")], (total, cta)   # sparse and literal-heavy ones are not
