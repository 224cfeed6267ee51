"""World-size-2 test (gloo, CPU) of the multi-rank host logic: shard plans partition the
replication ids, each rank computes its own replications (here with the oracle as a stand-in for
the GPU), the counters are all-reduced, and the result equals the single-rank run."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plans_partition():
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    for reps in (1, 7, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            ids = []
            for r in range(world):
                n, off = pkg.shard_plan(reps, world, r, "strong")
                ids += list(range(off, off + n))
            assert ids == list(range(reps))
            w = [pkg.shard_plan(reps, world, r, "weak") for r in range(world)]
            assert w == [(reps, r * reps) for r in range(world)]


def _worker(rank, world, port, reps, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, off = pkg.shard_plan(reps, world, rank, "strong")
    st = np.zeros((1, n), dtype=pkg.STATS_DTYPE)
    for i in range(n):
        res, _, _ = O.run_port(O.make_config(nUE=400, seed=9, rep=off + i), per_ue=False)
        for k in ("nSuccess", "preambleTxSum", "delaySum", "failCountSum", "continueFailed",
                  "collisionPreambles", "totalPreambleTxop"):
            st[0, i][k] = getattr(res, k)
        st[0, i]["updates"] = 400 * ((res.simTimeMs + 4) // 5)
    tot = pkg.allreduce_counters(pkg.local_counter_vector(st), dist)
    dist.barrier()
    if rank == 0:
        q.put(tot)
    dist.destroy_process_group()


def test_two_ranks_equal_one(oracle):
    import torch.multiprocessing as mp
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    reps = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reps, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = {k: 0 for k in pkg.COUNTER_KEYS}
    for i in range(reps):
        res, _, _ = oracle.run_port(oracle.make_config(nUE=400, seed=9, rep=i), per_ue=False)
        for k in pkg.COUNTER_KEYS:
            if k == "recordMoves":       # engine bookkeeping, not a reference counter: the stand-in leaves it 0
                continue
            exp[k] += 400 * ((res.simTimeMs + 4) // 5) if k == "updates" else getattr(res, k)
    assert tot["replications"] == reps
    for k in pkg.COUNTER_KEYS:
        assert tot[k] == exp[k], k
