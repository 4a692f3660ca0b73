"""Batch sharding of independent simulations over the GPUs of one box.

The PISO path has no cross-sample coupling, so the only multi-GPU logic is: which samples does a rank own, and how are
device timings / iteration statistics combined.  No data-path collective exists; the closure-network gradient
all-reduce of the training configs belongs to the training driver (torch.distributed all_reduce over NCCL)."""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous split of `total` samples; the first `total % world_size` ranks take one extra sample."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def local_batch(tensor, world_size=None, rank=None):
    """Slice [B, ...] down to the samples this rank owns."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    start, count = shard_bounds(tensor.shape[0], world_size, rank)
    return tensor[start:start + count]


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (device timings are reported as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_counts(local_value, device=None):
    """All ranks' integer values (e.g. samples processed) as a list, for whole-job aggregates."""
    t = torch.tensor([int(local_value)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [int(o.item()) for o in out]
    return [int(local_value)]
