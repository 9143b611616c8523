"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).

The reference is single-GPU (pl.Trainer(gpus=1): d3f/train_denoiser/train_denoiser.py:43-48).  The hot path
shards in two ways: sampling is batch-sharded with NO communication; training is data parallel
with ONE exchange step, the mean-allreduce of the flat gradient arena, issued per backward segment so it
overlaps the rest of backward.  These helpers are device-agnostic so the N>1 logic is covered by
world_size-2 gloo tests on CPU."""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Contiguous [start, end) of `total` units owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(x, world, rank):
    s, e = shard_range(x.shape[0], world, rank)
    return x[s:e]


def allreduce_bucket_(arena, start, end, group=None):
    """In-place mean over the data-parallel group of arena[start:end] (PL-DDP gradient semantics)."""
    view = arena[start:end]
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        view.div_(dist.get_world_size(group))
    return view


def broadcast_flat_(flat, src=0, group=None):
    """Make every replica start from rank `src`'s parameters."""
    dist.broadcast(flat, src=src, group=group)
    return flat


def gather_shards(local, world, group=None):
    """Concatenate batch shards from all ranks (used only to assemble sampler outputs for inspection)."""
    outs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(outs, local, group=group)
    return torch.cat(outs, dim=0)
