"""world_size-2 gloo test of the pair-sharding host logic (no GPU): each rank solves its block with the oracle
standing in for the solver, results are gathered and must equal the single-process result exactly."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from relativepose_b200.sharding import shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 9, 100):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rp_oracle
    from relativepose_b200 import synth
    from relativepose_b200.sharding import solve_sharded
    P = synth.shipped_params("suncg")
    recs = synth.make_batch(900, 5, 16)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    full = solve_sharded(recs, rp_oracle.Params(*P[0]), rp_oracle.solve_batch, rank, world, gather)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_equals_single():
    from oracle import rp_oracle
    from relativepose_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    P = synth.shipped_params("suncg")
    ref = rp_oracle.solve_batch(synth.make_batch(900, 5, 16), rp_oracle.Params(*P[0]))
    # scipy ARPACK draws a random start vector: the oracle is reproducible to ~1e-15, not bitwise
    assert np.abs(full - ref).max() <= 1e-10
