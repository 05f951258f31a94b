"""Multi-GPU plumbing for the only way this path shards: independent replicas.

One process per GPU; problem instances (seeds / KKT systems) are dealt round-robin to ranks, there is no
data-path collective, and the only communication is the max-over-ranks reduction of the step time plus a
gather of the per-rank results.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def assign(items, rank, world):
    """Instances owned by `rank`: round-robin, so every rank gets the same number +-1."""
    return [it for q, it in enumerate(items) if q % world == rank]


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_results(obj):
    """All ranks' result objects on every rank (list indexed by rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
