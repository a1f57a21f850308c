"""Data-parallel plumbing of the NCA step on one multi-GPU box.

The path shards by independent units — samples of the batch / pool, or independent frame streams — so there
is no state exchange.  The only collective is one all-reduce (sum) of the flattened weight gradients per
backward (6-13 k floats): NCCL over NVLink on the GPU box, gloo in the CPU tests.  Because the reference
L2-normalises every parameter's gradient before Adam (experiments.py:252-253, conditioned_trainer.py:134-136)
sum and mean give the same update, and every rank then applies the identical optimiser step (no broadcast).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n units owned by `rank`; sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed, rank):
    """Philox key of a rank: disjoint fire-mask streams across ranks for the same user seed."""
    return (int(seed) * 0x9E3779B97F4A7C15 + (rank + 1) * 0xD1B54A32D192ED03) & ((1 << 62) - 1)


def flat_view(tensors):
    """A zero-copy flat view over `tensors` when they sit back to back in one storage (the BPTT entry points write the weight
    gradients of a model that way, in parameter order), else None."""
    tensors = list(tensors)
    if not tensors or any(t is None for t in tensors):
        return None
    t0 = tensors[0]
    base, o = t0.untyped_storage().data_ptr(), t0.storage_offset()
    for t in tensors:
        if (not t.is_contiguous() or t.dtype != t0.dtype or t.device != t0.device or t.untyped_storage().data_ptr() != base
                or t.storage_offset() != o):
            return None
        o += t.numel()
    return t0.as_strided((o - t0.storage_offset(),), (1,), t0.storage_offset())


def flatten_grads(params, out=None):
    """Concatenate the .grad of every parameter (zeros where None) into one flat fp32 buffer."""
    params = list(params)
    n = sum(p.numel() for p in params)
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=params[0].device)
    o = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            out[o:o + k].zero_()
        else:
            out[o:o + k].copy_(p.grad.reshape(-1))
        o += k
    return out


def unflatten_grads(params, flat):
    o = 0
    for p in params:
        k = p.numel()
        g = flat[o:o + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += k


def allreduce_grads(params, group=None, flat=None):
    """Sum the weight gradients over all ranks in ONE collective; returns the flat buffer."""
    params = list(params)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    view = flat_view([p.grad for p in params]) if flat is None else None
    if view is not None:        # the gradients already are one buffer: reduce it in place (no flatten, no copy back)
        if multi:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        return view
    flat = flatten_grads(params, flat)
    if multi:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unflatten_grads(params, flat)
    return flat


def normalize_grads_(params, eps=1e-8):
    """p.grad /= ||p.grad|| + eps per parameter (experiments.py:252-253)."""
    for p in params:
        if p.grad is not None:
            p.grad.div_(p.grad.norm() + eps)


def max_over_ranks_ms(ms, device, group=None):
    """Device-timed milliseconds -> max over ranks (multi-GPU numbers are never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(ms)
    t = torch.tensor([float(ms)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
