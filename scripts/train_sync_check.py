"""2-GPU check of the Alg-2 training step (run under torchrun, one rank per GPU):

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_sync_check.py

(1) sync_bn=True: one optimizer step of FlowTrainer on N ranks, each with its own minibatch, must equal one
    single-process step on the concatenated batch (same initial weights, same Adam) -- main_algorithm_2.py:437-452 is a
    single process, so this is the N-GPU semantics that reproduces it;
(2) sync_bn=False (default): gradients are averaged, parameters stay identical across ranks, BatchNorm running
    statistics differ per rank until broadcast_flow;
(3) broadcast_flow leaves every rank with rank 0's parameters and buffers (one flat broadcast each for float / int).
Prints one JSON line on rank 0.
"""
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowstate_b200.normflows as NF          # noqa: E402
from flowstate_b200 import parallel            # noqa: E402
from flowstate_b200.drivers.training import FlowTrainer   # noqa: E402


def build(n, K, blocks, H, nb, bound, dev, seed=0):
    torch.manual_seed(seed)
    flows = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, blocks, H, list(range(2 * n)), num_bins=nb,
                                                             tail_bound=bound) for _ in range(K)]
    m = NF.NormalizingFlow(NF.Energy.UniformParticle(n, 2, bound, dev), flows).to(dev)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_((0.05 * torch.randn(p.shape, generator=g)).to(dev))
    return m


def max_diff(a, b):
    return max(float((x.detach().double() - y.detach().double()).abs().max()) for x, y in zip(a, b) if x.numel())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n, K, blocks, H, nb, bound, rows = 8, 3, 2, 128, 15, 8.0, 64
    g = torch.Generator().manual_seed(7)
    full = ((torch.rand(world * rows, 2 * n, generator=g) * 2 - 1) * bound * 0.9).to(dev)
    mine = full[rank * rows:(rank + 1) * rows].contiguous()
    out = {}

    # (1) synchronised statistics == single process on the union batch
    ref = build(n, K, blocks, H, nb, bound, dev).train()
    m = copy.deepcopy(ref).train()
    import flowstate_b200.drivers.training as T
    t_ref = FlowTrainer(ref, 1e-3, 1e-4, use_graph=False)
    keep, T._world = T._world, (lambda: 1)                     # single-process semantics for the reference step
    loss_ref = t_ref.step(full)
    T._world = keep
    t = FlowTrainer(m, 1e-3, 1e-4, use_graph=True, sync_bn=True)
    loss = t.step(mine)
    # gradients (the flat buckets), not parameters: Adam's first step is lr * g / (|g| + eps), which amplifies
    # rounding differences of near-zero gradients to +-lr
    gmax = float(t_ref.flat.abs().max())
    out["sync_bn_grad_diff_rel"] = float((t.flat - t_ref.flat).abs().max()) / gmax
    out["sync_bn_buffer_diff"] = max_diff([b for b in m.buffers() if b.is_floating_point()],
                                          [b for b in ref.buffers() if b.is_floating_point()])
    lt = torch.tensor([loss], device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    out["sync_bn_loss_diff"] = abs(float(lt) / world - loss_ref)

    # (2) default: averaged gradients, identical parameters, per-rank running statistics
    m2 = build(n, K, blocks, H, nb, bound, dev).train()
    t2 = FlowTrainer(m2, 1e-3, 1e-4, use_graph=True)
    for _ in range(2):
        t2.step(mine)
    flat = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["ddp_param_spread"] = float((hi - lo).abs().max())
    bufs = torch.cat([b.detach().reshape(-1).float() for b in m2.buffers() if b.is_floating_point()])
    lo, hi = bufs.clone(), bufs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["ddp_running_stat_spread_before_broadcast"] = float((hi - lo).abs().max())
    out["allreduce_calls"], out["allreduce_bytes"] = t2.allreduce_calls, t2.allreduce_bytes

    # (3) broadcast
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    parallel.broadcast_flow(m2, src=0)
    e0.record()
    nbytes = parallel.broadcast_flow(m2, src=0)
    e1.record()
    torch.cuda.synchronize()
    out["broadcast_bytes"], out["broadcast_ms"] = nbytes, e0.elapsed_time(e1)
    every = torch.cat([t_.detach().reshape(-1).double() for t_ in list(m2.parameters()) + list(m2.buffers())])
    lo, hi = every.clone(), every.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["spread_after_broadcast"] = float((hi - lo).abs().max())
    ok = (out["sync_bn_grad_diff_rel"] < 1e-4 and out["sync_bn_buffer_diff"] < 1e-5 and out["ddp_param_spread"] == 0.0
          and out["ddp_running_stat_spread_before_broadcast"] > 0.0 and out["spread_after_broadcast"] == 0.0)
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
