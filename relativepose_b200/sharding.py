"""Pair sharding across ranks: scan pairs are independent, so the only multi-GPU
structure is a partition of the pair list -- contiguous blocks, no collective on the data path.
The reference shards the same way with separate processes (--entrySplit, evaluation.py:59, datasets/SUNCG.py:68-69)."""
import numpy as np


def shard_bounds(num_pairs, rank, world_size):
    """Contiguous [lo, hi) block of pairs for `rank`; blocks differ in size by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(num_pairs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def solve_sharded(records, para, solve_fn, rank, world_size, gather_fn=None):
    """Solve this rank's block with `solve_fn(list_of_records, para) -> [b,4,4]`; when `gather_fn` is given
    (e.g. torch.distributed.all_gather_object) return the full [B,4,4] array on every rank."""
    lo, hi = shard_bounds(len(records), rank, world_size)
    local = solve_fn(records[lo:hi], para) if hi > lo else np.zeros([0, 4, 4])
    if gather_fn is None:
        return local
    parts = gather_fn((lo, np.asarray(local)))
    out = np.zeros([len(records), 4, 4])
    for plo, arr in parts:
        out[plo:plo + arr.shape[0]] = arr
    return out
