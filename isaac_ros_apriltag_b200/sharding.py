"""Frame-level data parallelism: frames are independent, so a batch shards across the GPUs of
one box with NO data-path collective -- one process, one detector handle and one CUDA stream per GPU.  torch.distributed
is used only for the launch plumbing (barrier, max-over-ranks timing, gathering the small per-frame results)."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous block partition: rank r owns [lo, hi).  Block sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def max_over_ranks(seconds, device=None):
    """Timing rule for every multi-GPU number: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(units_local, seconds_local, device=None):
    """units all ranks processed / max-over-ranks time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units_local / seconds_local
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / max_over_ranks(seconds_local, dev)


def run_sharded(detect_fn, frames, rank=None, world=None):
    """detect_fn(frames_slice) -> list (per frame) of results.  Every rank processes its block; rank 0 receives the
    concatenated per-frame results in frame order (gather of small host objects, not a data-path collective)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(len(frames), rank, world)
    local = detect_fn(frames[lo:hi]) if hi > lo else []
    if world == 1:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    out = []
    for part in gathered:
        out.extend(part)
    return out
