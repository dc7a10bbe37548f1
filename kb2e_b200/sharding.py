"""Multi-GPU plumbing for the ranking pass: one process per GPU, test triples sharded by contiguous
range, the four int64 sums added across ranks with one all-reduce (NCCL on GPUs, gloo in the CPU tests).
The data path itself has no collective: tables and the filter set are replicated."""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of rank's share; the union over ranks is exactly [0, n_items)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_items * rank // world, n_items * (rank + 1) // world


def allreduce_sums(sums, device=None):
    """Sum an int64 vector over the default torch.distributed group (no-op without one)."""
    import torch
    import torch.distributed as dist

    a = np.asarray(sums, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return a
    t = torch.from_numpy(a.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def merge_metrics(sums, n_queries):
    """The reference's four printed numbers (common/evaluation.cpp:247-250) from the global sums."""
    s = np.asarray(sums, dtype=np.float64)
    return {"raw_mean_rank": s[0] / n_queries, "filtered_mean_rank": s[1] / n_queries,
            "raw_hits10": s[2] / n_queries, "filtered_hits10": s[3] / n_queries}
