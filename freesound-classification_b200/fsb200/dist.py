"""Data-parallel plumbing (one process per GPU, `torch.distributed`): the reference has no distributed code, so
this is new.  The path shards on the batch (independent clips, per-rank BatchNorm statistics -- SURVEY.md 8e) and
has ONE exchange step: a SUM all-reduce of the flat float32 gradient; the division by the world size is folded into
the Adam kernel (`grad_scale`).  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [begin, end) slice of `n_items` independent units (clips / buckets) for `rank`."""
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_gradients(grads, flat=None):
    """SUM all-reduce of the gradients of one optimiser step.

    flat: the plan's persistent flat gradient buffer when every tensor in `grads` is a view of it (the normal
    case: ONE collective, in place); otherwise the gradients are packed into a temporary flat tensor, reduced with one
    collective and copied back.  Returns the factor the optimiser must apply to the summed gradient (1 / world)."""
    rank, world_size = world()
    if world_size == 1:
        return 1.0
    grads = list(grads)
    aliased = (flat is not None and len(grads) > 0 and grads[0].data_ptr() == flat.data_ptr()
               and sum(g.numel() for g in grads) == flat.numel())
    if aliased:
        dist.all_reduce(flat)
    else:
        packed = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(packed)
        off = 0
        for g in grads:
            g.copy_(packed[off:off + g.numel()].view_as(g))
            off += g.numel()
    return 1.0 / world_size
