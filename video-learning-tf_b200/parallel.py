"""Data-parallel plumbing (one process per GPU; the reference itself is single-process).

Clips are independent through the whole forward pass, so the batch is sharded contiguously across ranks (whole
videos stay on one rank) and the only exchange is the gradient all-reduce (+ two scalars).  torch.distributed is the
transport: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(num_items, rank, world):
    """Contiguous [lo, hi) of items (videos) owned by `rank`; the first `num_items % world` ranks get one more."""
    base, rem = divmod(num_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_grad_scale(local_clips, world):
    """d(loss)/d(logits) scale so that SUMMING rank gradients gives the gradient of the global-batch mean loss
    (train.py:123 reduce_mean over the whole batch); requires equal `local_clips` on all ranks."""
    return 1.0 / (local_clips * world)


def allreduce_gradients(flat_grads, scalars=None, group=None):
    """Sum the flat gradient arena (and the [loss_mean, correct] scalars) over ranks, in place."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    if scalars is not None:
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM, group=group)


def allreduce_async(flat_slice, group=None):
    """Start summing a slice of the gradient arena over ranks; returns a handle for `wait` (None when world == 1).
    NCCL orders the collective after the work already enqueued on the current stream and runs it on its own
    stream, so kernels enqueued afterwards overlap with the transfer."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=group, async_op=True)


def wait(handle):
    """Make the current stream wait for an `allreduce_async` handle."""
    if handle is not None:
        handle.wait()


def gather_logits(local_logits, group=None):
    """Validation: all ranks' clip logits on every rank, in rank order (contiguous shards -> original order)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_logits
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_logits.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_logits.shape[0]], dtype=torch.int64, device=local_logits.device),
                    group=group)
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros(mx, local_logits.shape[1], dtype=local_logits.dtype, device=local_logits.device)
    pad[:local_logits.shape[0]] = local_logits
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:int(s.item())] for o, s in zip(outs, sizes)], dim=0)
