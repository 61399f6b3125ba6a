"""-m gpu: batched LZ4 block compression (k_lz4_encode) against the reference.

Replaces LZ4_compress_fast at reference compression.c:70-72.  Three-part contract
(BASELINE.json north_star): every GPU-written block is a standard LZ4 block that the
reference's own cryo_decompress (liblz4) restores byte-identically; its size stays
within the stated tolerance of the reference's size at the same lz4_acceleration;
and the GPU decoder reads it back too.
"""
import numpy as np
import pytest

from pg_cryogen_b200 import COMP_LZ4, CRYO_BLCKSZ, compress_bound
from pg_cryogen_b200 import blockgen as bg

from gpu_util import decode_device, encode_device

pytestmark = pytest.mark.gpu

TOL = 1.10                  # gpu_csize <= 1.10 x reference + SLACK at EVERY acceleration (DESIGN.md section 1;
SLACK = 64                  # measured worst case per acceleration: profiles/r02_ratio_sweep.txt)


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
    rnd = np.frombuffer(bg.rand_bytes(99, CRYO_BLCKSZ).tobytes(), dtype=np.uint8)
    blocks.append(rnd.copy())
    tags.append("incompressible")
    return np.stack(blocks), tags


@pytest.mark.parametrize("accel", [0, 1, 2, 5, 10, 25, 50])
def test_lz4_encode_roundtrip_and_ratio(gpu, oracle_ref, accel):
    blocks, tags = _blocks()
    comp, st = encode_device(gpu, COMP_LZ4, accel, blocks)
    assert (st == 0).all(), st
    ref_comp, ref_sizes, _ = oracle_ref.compress(COMP_LZ4, accel, blocks, nthreads=8)
    back, ok, _ = oracle_ref.decompress([COMP_LZ4] * len(comp), *oracle_ref.pack(comp))
    assert ok.all(), "the reference's LZ4_decompress_safe rejected a GPU-written block"
    assert np.array_equal(back, blocks)
    for i, c in enumerate(comp):
        assert len(c) <= compress_bound(COMP_LZ4), tags[i]
        assert len(c) <= TOL * ref_sizes[i] + SLACK, (tags[i], accel, len(c), int(ref_sizes[i]))
    out, osz, dst = decode_device(gpu, COMP_LZ4, comp)
    assert (dst == 0).all() and np.array_equal(out, blocks)


def test_lz4_encode_acceleration_zero_means_one(gpu):
    """liblz4 treats acceleration < 1 as 1; the GUC allows 0 (compression.c:41)."""
    blk = bg.make_block("M", "lowcard", 4)[None]
    a, _ = encode_device(gpu, COMP_LZ4, 0, blk)
    b, _ = encode_device(gpu, COMP_LZ4, 1, blk)
    assert np.array_equal(a[0], b[0])


def test_lz4_encode_small_block_sizes(gpu, oracle_ref):
    """block_size is a runtime parameter (64 KiB blocks)."""
    import ctypes as C
    lz4 = C.CDLL("liblz4.so.1")
    lz4.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    src = bg.make_block("D", "lowcard", 2)
    for bs in (65536, 4096, 1024, 64, 16):
        blocks = src[: bs * 4].reshape(4, bs)
        comp, st = encode_device(gpu, COMP_LZ4, 1, blocks, block_size=bs)
        assert (st == 0).all()
        for i, c in enumerate(comp):
            out = np.zeros(bs, dtype=np.uint8)
            n = lz4.LZ4_decompress_safe(c.ctypes.data, out.ctypes.data, len(c), bs)
            assert n == bs and np.array_equal(out, blocks[i]), (bs, i, n)


def test_lz4_encode_host_api_and_shim_contract(gpu, oracle_ref):
    blocks = np.stack([bg.make_block(k, "hex", 40 + i) for i, k in enumerate("SMD")])
    comp, st = gpu.compress_host(COMP_LZ4, 1, blocks)
    assert (st == 0).all()
    back, ok, _ = oracle_ref.decompress([COMP_LZ4] * 3, *oracle_ref.pack(comp))
    assert ok.all() and np.array_equal(back, blocks)
