"""Multi-GPU plumbing for bench.py: one process per GPU, independent map replicas (no data-path collective).

The per-frame path of the reference is single-GPU.  Until the z-slab sharded 1024^3
configuration exists, N GPUs run N independent maps on different sensor streams; the only communication
is the timing reduction below (torch.distributed, NCCL on GPUs / gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def replica_seed(base_seed, rank):
    """Each replica maps its own seeded world / trajectory."""
    return int(base_seed) + 1000 * int(rank)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(ms_per_step_local, device="cpu"):
    """Whole-job frames/s: every rank processes one frame per step; time is the max over ranks."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(ms_per_step_local, device)
    return world * 1000.0 / ms, ms
