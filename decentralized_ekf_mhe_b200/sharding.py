"""Instance sharding across the GPUs of one box: rank r of R owns the contiguous range
[r*n/R, (r+1)*n/R).  Instances are independent (DecentralEst.hpp:105-291 keeps all state private to
the object), so there is NO collective on the hot path; torch.distributed is used only for the
barrier around the timed region and the max-over-ranks reduction of the device time."""
import os


def shard_range(n_total, rank, world_size):
    """Contiguous instance range [lo, hi) owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = (n_total * rank) // world_size
    hi = (n_total * (rank + 1)) // world_size
    return lo, hi


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def max_over_ranks(value, device=None, group=None):
    """All-reduce MAX of a python float (device timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device=None, group=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
