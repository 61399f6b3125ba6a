"""Block-range sharding of a batch of cryo blocks over ranks / GPUs.

Cryo blocks are compressed and decompressed independently (reference compression.c:70-72, :84,
:102-104, :116 take one block, no dictionary, no shared state), so a batch is split into
contiguous block ranges, one per GPU, with no collective on the data path.
The same split is used by the C ABI (`cryogpu_*_host_multi`, run_sharded in cryogpu.cu) and by
`bench.py --gpus N` (one rank per GPU); the only communication is the timing barrier and the
max-over-ranks of the elapsed time.
"""
from __future__ import annotations


def block_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the blocks rank `rank` of `world` owns; identical to run_sharded()."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank/world out of range")
    return n * rank // world, n * (rank + 1) // world


def rank_block_seed_offset(rank: int, blocks_per_rank: int) -> int:
    """Weak scaling: rank r decompresses its own table whose block seeds start here."""
    return rank * blocks_per_rank


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a per-rank scalar (elapsed time); identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank: int, world: int, seconds_max: float) -> float:
    """Units all ranks processed divided by the slowest rank's time."""
    return world * units_per_rank / seconds_max
