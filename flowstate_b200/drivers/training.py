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
  * a fused multi-tensor Adam (torch.optim.Adam(fused=True)); a new optimizer per call of `fresh_optimizer`, like the
    reference's new Adam every cycle (main_algorithm_2.py:440);
  * the reference's "skip the step when the loss is NaN / Inf" decision taken collectively over the ranks;
  * optionally (sync_bn=True, N > 1 GPUs) train-mode BatchNorm statistics taken over the union of the ranks' batches
    (torch.nn.SyncBatchNorm, same state_dict keys): the N-GPU step then equals the single-process step on the
    concatenated batch; without it every rank normalises with its own batch (DDP's default semantics)
    and rank 0's running statistics are broadcast before sampling.  The synchronised pass runs eagerly (its
    collectives sit inside the forward and backward passes).
Parameters that never receive a gradient (PeriodicFeaturesElementwise.weights) keep grad = None, so
Adam skips them exactly like the reference's.
"""
import torch
import torch.distributed as dist

from . import _train_native


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


class FlowTrainer:
    def __init__(self, model, lr, weight_decay=0.0, alpha=1.0, reverse_batch=256, use_graph=True, sync_bn=False,
                 native=True):
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
        total = sum(p.numel() for p in self.trainable)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        off = 0
        for p in self.trainable:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def fresh_optimizer(self):
        """A new Adam (zero moments, step 0) over every parameter, as the reference creates one per training cycle
        (main_algorithm_2.py:440).  After the first cycle the existing optimizer's state is zeroed in place - the same
        state as a newly built one, in three multi-tensor launches instead of ~1.5 k allocations and fills."""
        self._native_checked = False
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

    def step(self, batch):
        """One optimizer step on `batch`.  Returns the loss as a Python float, or None when the step was skipped
        (fewer than two rows, or a non-finite loss on any rank)."""
        if self.opt is None:
            self.fresh_optimizer()
        world = _world()
        usable = batch.shape[0] >= 2                           # BatchNorm needs two rows
        if usable and self.flat is None:
            self._prepare(batch)
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
