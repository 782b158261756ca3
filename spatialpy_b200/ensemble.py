"""Ensemble sharding over GPUs: trajectories are independent units (the reference runs them as separate processes with
seed+k, spatialpy/solvers/solver.py:547-605), so the multi-GPU path has no data-path collective — only the
bookkeeping below and a max/sum over ranks for timing and counters."""


def shard_trajectories(number_of_trajectories, rank, world_size):
    """Trajectory indices owned by `rank`: k -> rank k mod world_size.  Trajectory k always uses seed + k, whatever the
    world size, so an ensemble is reproducible across GPU counts."""
    return list(range(rank, number_of_trajectories, world_size))


def reduce_scalar(value, op="max", device=None):
    """max / sum of a python float over the ranks of the default torch.distributed group (no-op when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())
