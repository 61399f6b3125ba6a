"""-m gpu: tuple-level work on device-resident decoded blocks against the item walk of the
reference's sequential scan (oracle/cryo_pages.c cryo_oracle_block_tuple_stats, pinned to the reference's
storage.c in tests/test_pages.py)."""
import numpy as np
import pytest
import torch

from oracle import pages as opg
from pg_cryogen_b200 import CRYO_BLCKSZ, blockgen as bg

pytestmark = pytest.mark.gpu


def test_count_pushdown_matches_the_scan(gpu, oracle_ref):
    kinds = [("S", "hex"), ("M", "lowcard"), ("D", "hex"), ("D", "lowcard"), ("S", "random")]
    blocks = [bg.make_block(k, p, 40 + i) for i, (k, p) in enumerate(kinds)]
    blocks.append(bg.regression_block(1, 290))
    empty = np.zeros(CRYO_BLCKSZ, dtype=np.uint8)
    empty[0] = 8                                        # cryo_init_page: lower = 8, upper = 1 MiB
    empty[4:8] = np.frombuffer(np.uint32(CRYO_BLCKSZ).tobytes(), dtype=np.uint8)
    blocks.append(empty)
    bad = blocks[0].copy()
    bad[8 + 8 * 3: 8 + 8 * 3 + 4] = 0xFF                # item 4 points outside the block
    blocks.append(bad)
    blocks = np.stack(blocks)
    n = blocks.shape[0]
    methods = [i & 1 for i in range(n)]
    comp = [oracle_ref.compress(m, 1, b)[0][0] for m, b in zip(methods, blocks)]
    comp.append(comp[2][:-9])                           # a truncated stream: its status is the decoder's
    methods.append(methods[2])
    nt, by, st = gpu.decompress_count_host(methods, comp)
    for i in range(n):
        want = opg.block_tuple_stats(blocks[i])
        assert (int(nt[i]), int(by[i])) == want[:2], i
        assert st[i] == (0 if want[2] else 4), (i, st[i])
    assert st[n] != 0 and nt[n] == 0
    h2d, d2h = gpu.last_transfer_bytes()
    assert d2h < 1024                                   # no block came back


def test_tuple_stats_device(gpu):
    dev = torch.device("cuda", gpu.device)
    blocks = np.stack([bg.make_block("S", "hex", 50 + i) for i in range(37)])
    d = torch.from_numpy(blocks).to(dev)
    nt = torch.zeros(37, dtype=torch.int32, device=dev)
    by = torch.zeros(37, dtype=torch.int64, device=dev)
    ok = torch.zeros(37, dtype=torch.int32, device=dev)
    rc = gpu.lib.cryogpu_tuple_stats_device(gpu.handle, 37, d.data_ptr(), CRYO_BLCKSZ, CRYO_BLCKSZ, None, nt.data_ptr(),
                                            by.data_ptr(), ok.data_ptr(), torch.cuda.current_stream(dev).cuda_stream or 1)
    assert rc == 0
    torch.cuda.synchronize(dev)
    for i in range(37):
        assert (int(nt[i]), int(by[i]), int(ok[i])) == opg.block_tuple_stats(blocks[i])
