"""Sharding of replications over ranks (one process per GPU) and the one small all-reduce of the
per-replication counters at the end.  Replications are keyed by (seed, replication id) in the draw
tape, so a shard is just a contiguous range of replication ids: rank r runs
[rep_offset, rep_offset + reps_local) and no rank needs anything from another until the counters
are summed.  Reference: the seed loop and the nUE sweep are independent iterations
(RandomAccessWithNOMA.c:216,221)."""
import numpy as np

COUNTER_KEYS = ("updates", "nSuccess", "preambleTxSum", "delaySum", "failCountSum", "continueFailed",
                "collisionPreambles", "totalPreambleTxop", "recordMoves")


def shard_plan(reps, world, rank, scaling="weak"):
    """-> (reps_local, rep_offset).  weak: every rank runs `reps` (ids rank*reps ...);
    strong: `reps` in total, split as evenly as possible, lower ranks take the remainder."""
    if scaling == "weak":
        return reps, rank * reps
    base, rem = divmod(reps, world)
    local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return local, offset


def local_counter_vector(stats_all):
    """stats_all: structured array (points, reps) of ra_stats -> int64 vector [len(COUNTER_KEYS)+1]."""
    v = [int(stats_all[k].sum()) for k in COUNTER_KEYS]
    v.append(int(stats_all.size))
    return np.asarray(v, dtype=np.int64)


def allreduce_counters(vec, dist=None, device=None):
    """Sum the counter vector over ranks (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    t = torch.as_tensor(vec, dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    d = {k: int(out[i]) for i, k in enumerate(COUNTER_KEYS)}
    d["replications"] = int(out[-1])
    return d
