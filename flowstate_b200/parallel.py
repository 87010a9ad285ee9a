"""Multi-GPU plumbing: independent chains shard across ranks, flow weights are
broadcast, training gradients are all-reduced.

The reference has no distributed code at all; the hook points are
its driver loops: weights change at hybrid_NF_MCMC/main_algorithm_2.py:440-452
(optimizer step) and are consumed at :476-482, 534-548 (sampling / global moves).
One process per GPU over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
Chains never interact, so sampling itself needs no collective.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block of chain ids owned by `rank`: (start, count).  Per-chain RNG streams
    are keyed by the GLOBAL chain id, so results do not depend on `world`."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def _flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors]) if tensors else torch.empty(0)


def _broadcast_flat(tensors, src, group):
    """cat -> ONE broadcast -> multi-tensor scatter back (a handful of kernels however many tensors there are)."""
    if not tensors:
        return 0
    dtype = tensors[0].dtype
    flat = torch.cat([t.detach().reshape(-1).to(dtype) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    views = [v.view(t.shape) for v, t in zip(flat.split([t.numel() for t in tensors]), tensors)]
    with torch.no_grad():
        same = [i for i, t in enumerate(tensors) if t.dtype == dtype]
        torch._foreach_copy_([tensors[i].detach() for i in same], [views[i] for i in same])
        for i, t in enumerate(tensors):
            if t.dtype != dtype:
                t.copy_(views[i].to(t.dtype))
    return flat.numel() * flat.element_size()


def broadcast_flow(model, src=0, group=None, buffers_only=False):
    """Every parameter and buffer (BatchNorm statistics and their integer batch counters included) from `src`: one
    flat broadcast for the floating-point tensors, one for the integer buffers; call after training on rank `src` /
    before sampling.  buffers_only: the parameters are already identical on every rank (they started identical and
    every optimizer step used the all-reduced gradient bucket), only the per-rank BatchNorm running statistics need
    rank `src`'s values.  Returns the bytes broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    ts = [t for t in ([] if buffers_only else list(model.parameters())) + list(model.buffers())
          if t.is_floating_point()]
    ints = [t for t in model.buffers() if not t.is_floating_point()]
    nbytes = _broadcast_flat(ts, src, group) + _broadcast_flat(ints, src, group)
    if hasattr(model, "repack"):
        model.repack()
    return nbytes


def allreduce_gradients(model, group=None, average=True):
    """Sum (or average) gradients over ranks as ONE flat bucket; call between loss.backward()
    and optimizer.step() (main_algorithm_2.py:450-451)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    ps = [p for p in model.parameters() if p.grad is not None]
    if not ps:
        return 0
    flat = _flat([p.grad for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    torch._foreach_copy_([p.grad for p in ps],
                         [v.view(p.shape) for v, p in zip(flat.split([p.numel() for p in ps]), ps)])
    return flat.numel() * 4


def allreduce_counters(*tensors, group=None):
    """Global sums of int64 acceptance counters (only needed when statistics are reported)."""
    out = [t.sum().reshape(1).clone() for t in tensors]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        flat = torch.cat(out)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out = [flat[i:i + 1] for i in range(len(out))]
    return [int(o.item()) for o in out]
