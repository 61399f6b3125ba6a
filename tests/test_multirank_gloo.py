"""world_size-2 gloo test of the N > 1 path on CPU (not gpu): block-range sharding with no
data-path collective, max-over-ranks timing, whole-job aggregation -- the host logic of
`bench.py --gpus N` and of cryogpu_*_host_multi.  The per-rank "codec" here is
the oracle (this is a test of the sharding logic, not of the kernels)."""
import json
import os
import subprocess
import sys

import pytest

from pg_cryogen_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import hashlib, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["CRYO_ROOT"])
from pg_cryogen_b200 import shard, blockgen as bg
from oracle import ref, port

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 7                                             # blocks in the batch (odd: uneven ranges)
lo, hi = shard.block_range(N, rank, world)
t0 = time.perf_counter()
digests = {}
for b in range(lo, hi):
    blk = bg.make_block("S", "hex", b)
    method = b % 2                                # one batch mixes lz4 and zstd (storage.h:64)
    comp = ref.compress(method, 1, blk)[0][0]
    n, out = (port.lz4_decode if method == 0 else port.zstd_decode)(comp)[:2]
    assert n == 1 << 20 and np.array_equal(out, blk)
    digests[b] = hashlib.sha256(out.tobytes()).hexdigest()
elapsed = time.perf_counter() - t0 + 0.01 * rank  # make the slowest rank predictable
dist.barrier()
tmax = shard.max_over_ranks(elapsed)
gathered = [None] * world
dist.all_gather_object(gathered, {"rank": rank, "range": [lo, hi], "digests": digests, "t": elapsed})
if rank == 0:
    print(json.dumps({"tmax": tmax, "ranks": gathered,
                      "rate": shard.whole_job_rate(N // world, world, tmax)}))
dist.destroy_process_group()
'''


def test_block_range_is_a_partition():
    for n in (0, 1, 2, 7, 64, 3449):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.block_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.block_range(4, 2, 2)


def test_two_ranks_gloo_shard_without_collective(oracle_ref, oracle_port, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CRYO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
        env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    ranks = sorted(res["ranks"], key=lambda r: r["rank"])
    assert [tuple(r["range"]) for r in ranks] == [(0, 3), (3, 7)]
    seen = {}
    for r in ranks:
        for b, d in r["digests"].items():
            assert b not in seen, "a block was processed by two ranks"
            seen[b] = d
    assert sorted(int(b) for b in seen) == list(range(7))
    assert abs(res["tmax"] - max(r["t"] for r in ranks)) < 1e-9      # max over ranks, not rank 0's
    assert res["rate"] == pytest.approx(2 * 3 / res["tmax"])
