"""bench.py's gate (not gpu): the per-block digest is the same function in numpy (host side: the table as built,
the e2e destination) and in torch (device side), it sees moved, swapped and flipped bytes, and block-range
sharding covers every block exactly once."""
import numpy as np
import torch

import bench
from pg_cryogen_b200 import shard


def test_digest_is_the_same_in_numpy_and_torch_and_sees_small_changes():
    rng = np.random.default_rng(1)
    blocks = rng.integers(0, 256, size=(3, bench.CRYO_BLCKSZ), dtype=np.uint8)
    blocks[1, 5000:900000] = 0                      # a sparse block
    want = bench.digest_np(blocks)
    got = bench.digest_torch(torch.from_numpy(blocks.copy())).numpy()
    assert np.array_equal(want, got)
    for change in ("flip", "swap", "shift"):
        b = blocks.copy()
        if change == "flip":
            b[1, 123456] ^= 1
        elif change == "swap":
            b[0, 8:16], b[0, 16:24] = blocks[0, 16:24].copy(), blocks[0, 8:16].copy()
        else:
            b[2, 1000:2000] = np.roll(blocks[2, 1000:2000], 8)
        d = bench.digest_np(b)
        assert not np.array_equal(d, want), change
        assert (d != want).any(axis=1).sum() == 1   # only the block that changed


def test_block_ranges_cover_the_job_once():
    for n in (1, 7, 3449, 10240):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.block_range(n, r, world)
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
