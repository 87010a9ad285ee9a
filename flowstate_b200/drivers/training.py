"""Training steps of the flow for the hybrid drivers (Algorithm 1 pre-training, Algorithm 2 per-cycle updates).

The reference trains with plain eager autograd, one optimizer step per minibatch
(hybrid_NF_MCMC/main_algorithm_1.py:297-320, main_algorithm_2.py:437-452):

    optimizer.zero_grad(); loss = ALPHA * forward_kld(batch) + (1 - ALPHA) * reverse_kld(B)[0]
    if loss is finite: loss.backward(); optimizer.step()

Here the same step runs with
  * the forward-KL loss and all its gradients from the hand-written training kernels (fs_train_forward_kld,
    drivers/_train_native.py: ~36 launches for the whole flow, gradients written straight into the flat bucket) whenever
    the flow has that path (`native`); otherwise
  * the forward + backward pass captured once per batch shape in a CUDA graph and replayed
    (the autograd pass is ~5 k small kernels for the K = 23 flow: launch-bound when issued eagerly);
  * all gradients living in ONE flat float32 buffer (every p.grad is a view into it), so the multi-GPU gradient
    all-reduce is a single NCCL call on that buffer (SUM, then divided by the world size) with no flatten / unflatten
    copies - the hook point is between backward and optimizer.step (main_algorithm_2.py:450-451);
  * on CUDA, the parameters and both Adam moments in flat buffers as well (every Parameter's .data is a view, names and
    state_dict untouched), so optimizer.step() is fs_adam_step: two launches, the same update as torch.optim.Adam
    (tests/test_gpu_train.py), with the reference's "skip the step when the loss is NaN / Inf" decided on the device -
    the host does not wait for the loss (`step(batch, sync=False)`); on CPU a torch.optim.Adam.  A new optimizer per
    call of `fresh_optimizer` (zeroed moments and step count), like the reference's new Adam every cycle
    (main_algorithm_2.py:440);
  * the skip decision taken collectively over the ranks (MAX all-reduce of the flag);
  * optionally (sync_bn=True, N > 1 GPUs) train-mode BatchNorm statistics taken over the union of the ranks' batches
    (torch.nn.SyncBatchNorm, same state_dict keys): the N-GPU step then equals the single-process step on the
    concatenated batch; without it every rank normalises with its own batch (DDP's default semantics)
    and rank 0's running statistics are broadcast before sampling.  The synchronised pass runs eagerly (its
    collectives sit inside the forward and backward passes).
Parameters that never receive a gradient (PeriodicFeaturesElementwise.weights) keep grad = None, so
Adam skips them exactly like the reference's.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _train_native
from .. import _lib


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


class FlowTrainer:
    def __init__(self, model, lr, weight_decay=0.0, alpha=1.0, reverse_batch=256, use_graph=True, sync_bn=False,
                 native=True, native_adam=True):
        self.sync_bn = bool(sync_bn) and _world() > 1
        if self.sync_bn:                                       # children are replaced in place; parameters are kept
            torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            use_graph = False
        self.model = model
        self.lr, self.weight_decay, self.alpha, self.reverse_batch = lr, weight_decay, alpha, reverse_batch
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.device = self.params[0].device
        self.use_graph = bool(use_graph) and self.device.type == "cuda" and alpha >= 1.0
        # hand-written forward + backward (pure forward-KL steps on flows that have the path)
        self.use_native = bool(native) and self.device.type == "cuda" and alpha >= 1.0 and not self.sync_bn \
            and _train_native.supported(model)
        self.native = None
        self._native_checked = False
        self._opt_kw = None
        # flat Adam (fs_adam_step): parameters, moments and step state in flat device buffers
        self.native_adam = bool(native_adam) and self.device.type == "cuda"
        self.flat_p = self.exp_avg = self.exp_avg_sq = self.adam_state = None
        self.flat = None                 # flat gradient bucket
        self.trainable = None            # parameters that receive gradients
        self.graphs = {}                 # batch rows -> (graph, static input, static loss)
        self.opt = None
        self.allreduce_bytes = 0
        self.allreduce_calls = 0

    # -- setup --------------------------------------------------------------
    def _loss(self, batch):
        loss = self.model.forward_kld(batch)
        if self.alpha < 1.0:                                   # main_algorithm_2.py:446-448
            energy_loss, _ = self.model.reverse_kld(self.reverse_batch)
            loss = self.alpha * loss + (1.0 - self.alpha) * energy_loss
        return loss

    def _prepare(self, batch):
        """One eager pass to find the parameters that get gradients, then the flat bucket with the grads as views."""
        for p in self.params:
            p.grad = None
        side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        if side is not None:                                   # off the default stream, like the graph warm-up below
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._loss(batch).backward()
            torch.cuda.current_stream(self.device).wait_stream(side)
        else:
            self._loss(batch).backward()
        self.trainable = [p for p in self.params if p.grad is not None]
        # every tensor starts on a 16-byte boundary of the flat buffers (the training kernels load float4)
        offs, total = [], 0
        for p in self.trainable:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        for p, off in zip(self.trainable, offs):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        if self.native_adam:
            self.flat_p = torch.zeros(total, dtype=torch.float32, device=self.device)
            for p, off in zip(self.trainable, offs):
                dst = self.flat_p[off:off + p.numel()].view_as(p)
                dst.copy_(p.data)
                p.data = dst                                   # same Parameter object, same state_dict entry
            self.exp_avg = torch.zeros_like(self.flat_p)
            self.exp_avg_sq = torch.zeros_like(self.flat_p)
            self.adam_state = torch.zeros(4, dtype=torch.float32, device=self.device)

    def fresh_optimizer(self):
        """A new Adam (zero moments, step 0) over every parameter, as the reference creates one per training cycle
        (main_algorithm_2.py:440).  After the first cycle the existing optimizer's state is zeroed in place - the same
        state as a newly built one, in three multi-tensor launches instead of ~1.5 k allocations and fills."""
        self._native_checked = False
        if self.native_adam:
            if self.adam_state is not None:
                self.exp_avg.zero_()
                self.exp_avg_sq.zero_()
                self.adam_state.zero_()
            self.opt = "fs_adam_step"
            return self.opt
        if self.opt is not None and self._opt_kw == (self.lr, self.weight_decay):
            bufs = [v for st in self.opt.state.values() for v in st.values() if torch.is_tensor(v)]
            if bufs:
                by_kind = {}
                for v in bufs:
                    by_kind.setdefault((v.dtype, v.device), []).append(v)
                for vs in by_kind.values():
                    torch._foreach_zero_(vs)
            return self.opt
        kw = dict(lr=self.lr, weight_decay=self.weight_decay)
        if self.device.type == "cuda":
            kw["fused"] = True
        self.opt = torch.optim.Adam(self.params, **kw)
        self._opt_kw = (self.lr, self.weight_decay)
        return self.opt

    # -- one minibatch ----------------------------------------------------------
    def _forward_backward(self, batch):
        """Fills the flat gradient bucket, returns the (device) loss."""
        rows = batch.shape[0]
        if self.use_native and self.model.training:
            if self.native is None or (not self._native_checked and not self.native.pointers_valid()):
                try:
                    self.native = _train_native.NativeForwardKL(self.model)
                except _train_native._lib.FlowStateError:      # FS_ERR_UNSUPPORTED (e.g. odd N): autograd path
                    self.use_native = False
                    self.native = None
            if self.native is not None:
                self._native_checked = True                    # re-validated once per optimizer (training cycle)
                return self.native.step(batch)[0]
        if not self.use_graph:
            self.flat.zero_()
            loss = self._loss(batch)
            if torch.isfinite(loss):
                loss.backward()
            return loss.detach()
        entry = self.graphs.get(rows)
        if entry is None:
            static_in = batch.clone()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                      # warm-up on a side stream (torch.cuda.graphs recipe)
                for _ in range(2):
                    self.flat.zero_()
                    self._loss(static_in).backward()
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            pool = next(iter(self.graphs.values()))[0].pool() if self.graphs else None
            with torch.cuda.graph(g, pool=pool):
                self.flat.zero_()
                loss = self._loss(static_in)
                loss.backward()                                # accumulates in place into the views of self.flat
                static_loss = loss.detach()                    # keeps the storage, drops the autograd graph (and with it
            del loss                                           # the AccumulateGrad nodes bound to the capture stream)
            entry = (g, static_in, static_loss)
            self.graphs[rows] = entry
        g, static_in, static_loss = entry
        static_in.copy_(batch)
        g.replay()
        return static_loss.detach()

    def step(self, batch, sync=True):
        """One optimizer step on `batch`.  sync=True returns the loss as a Python float, or None when the step was
        skipped (fewer than two rows, or a non-finite loss on any rank).  sync=False (flat Adam only) never waits for the
        device: it returns the loss as a 1-element device tensor (NaN for an unusable batch); a skipped step shows as
        a non-finite loss."""
        if self.opt is None:
            self.fresh_optimizer()
        world = _world()
        usable = batch.shape[0] >= 2                           # BatchNorm needs two rows
        if usable and self.flat is None:
            self._prepare(batch)
        if self.native_adam:
            return self._step_flat(batch, usable, world, sync)
        loss = self._forward_backward(batch) if usable else None
        bad = torch.zeros(1, device=self.device)
        if not usable:
            bad.fill_(1.0)
        else:
            bad = (~torch.isfinite(loss)).float().reshape(1)
        if world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if bad.item() > 0:                                     # every rank skips this step together
            return None
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)   # ONE collective on the flat bucket
            self.flat.div_(world)
            self.allreduce_bytes += self.flat.numel() * 4
            self.allreduce_calls += 1
        self.opt.step()
        return float(loss)

    def _step_flat(self, batch, usable, world, sync):
        if self.flat is None:                                  # nothing prepared yet and this batch cannot prepare it
            if world > 1:
                raise RuntimeError("FlowTrainer: the first minibatch of a rank needs at least two rows")
            return None if sync else torch.full((1,), float("nan"), device=self.device)
        loss = self._forward_backward(batch).reshape(1) if usable else None
        skip = None
        if world > 1 or not usable:
            skip = (~torch.isfinite(loss)).float() if usable else torch.ones(1, device=self.device)
            if world > 1:
                dist.all_reduce(skip, op=dist.ReduceOp.MAX)    # every rank skips this step together
        if world > 1:                                          # same collectives on every rank whatever the decision
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)   # ONE collective on the flat bucket
            self.flat.div_(world)
            self.allreduce_bytes += self.flat.numel() * 4
            self.allreduce_calls += 1
        _lib.check(_lib.lib().fs_adam_step(
            _lib.ptr(self.flat_p), _lib.ptr(self.flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
            int(self.flat_p.numel()), _lib.ptr(self.adam_state), _lib.ptr(skip), _lib.ptr(loss),
            float(self.lr), 0.9, 0.999, 1e-8, float(self.weight_decay), _lib.stream_ptr(self.device)))
        self._bump_versions()
        if not sync:
            return loss.clone() if usable else torch.full((1,), float("nan"), device=self.device)
        if self.adam_state[1].item() == 0:
            return None
        return float(loss)

    def _bump_versions(self):
        """fs_adam_step writes the parameters behind autograd's back; one in-place no-op on the flat buffer's views
        would cost a launch per tensor, so the model's inference pack is marked stale instead."""
        repack = getattr(self.model, "repack", None)
        if repack is not None:
            repack()
