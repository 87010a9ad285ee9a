"""Host side of fs_train_forward_kld: the forward-KL loss and ALL its gradients in ~36 kernel launches.

`NativeForwardKL(model, grads)` collects, per CircularCoupledRationalQuadraticSpline layer, the device pointers of the
parameters in the order include/flowstate_b200.h documents (module tree of NF/normflows/nets/resnet.py:7-104 and
flows/neural_spline/coupling.py:176-265), the pointers of their gradient tensors (views of the trainer's flat bucket)
and of the BatchNorm running statistics.  `step(batch)` then equals

    loss = model.forward_kld(batch); loss.backward()          (NF/normflows/core.py:88-108)

in training mode: gradients are written (not accumulated), running_mean / running_var / num_batches_tracked move as
torch.nn.BatchNorm1d moves them.  `supported(model)` says whether a flow has this path; the trainer keeps the autograd
path for everything else (reverse-KL mixing, SyncBatchNorm, dropout, odd particle numbers).
"""
import ctypes as C

import torch

from .. import _lib


def _layers(model):
    from ..normflows.flows import CircularCoupledRationalQuadraticSpline
    flows = list(getattr(model, "flows", []))
    if not flows or not all(isinstance(f, CircularCoupledRationalQuadraticSpline) for f in flows):
        return None
    return flows


def _tensors(flow):
    """(params, bn running buffers, num_batches_tracked buffers) of one layer in the C ABI's order."""
    c = flow.prqct
    net = c.transform_net
    ps = [net.initial_layer.weight, net.initial_layer.bias]
    rs, nbt = [], []
    for blk in net.blocks:
        for j in (0, 1):
            bn, lin = blk.batch_norm_layers[j], blk.linear_layers[j]
            ps += [bn.weight, bn.bias, lin.weight, lin.bias]
            rs += [bn.running_mean, bn.running_var]
            nbt.append(bn.num_batches_tracked)
    u = c.unconditional_transform
    ps += [net.final_layer.weight, net.final_layer.bias, u.unnormalized_widths, u.unnormalized_heights,
           u.unnormalized_derivatives]
    return ps, rs, nbt


def supported(model):
    flows = _layers(model)
    if flows is None:
        return False
    f0 = flows[0]
    net0 = f0.prqct.transform_net
    for f in flows:
        c, net = f.prqct, f.prqct.transform_net
        if not getattr(f, "fused_training", True):
            return False
        if not all(getattr(b, "use_batch_norm", False) and b.dropout.p == 0.0 for b in net.blocks):
            return False
        if any(type(bn) is not torch.nn.BatchNorm1d or bn.momentum is None or not bn.affine or not bn.track_running_stats
               for b in net.blocks for bn in b.batch_norm_layers):
            return False
        if (c.features != f0.prqct.features or c.num_bins != f0.prqct.num_bins or len(net.blocks) != len(net0.blocks)
                or net.hidden_features != net0.hidden_features or float(c.tail_bound) != float(f0.prqct.tail_bound)
                or not torch.equal(c.transform_features, f0.prqct.transform_features)):
            return False
        if torch.is_tensor(net.preprocessing.scale) or net.preprocessing.apply_bias:
            return False
    p = next(model.parameters())
    return p.is_cuda and p.dtype == torch.float32


class NativeForwardKL:
    def __init__(self, model):
        flows = _layers(model)
        self.model = model
        c0 = flows[0].prqct
        net0 = c0.transform_net
        self.K, self.N, self.H = len(flows), len(c0.transform_features), net0.hidden_features
        self.n_blocks, self.nb = len(net0.blocks), c0.num_bins
        params, running, self.nbt = [], [], []
        for f in flows:
            ps, rs, nbt = _tensors(f)
            params += ps
            running += rs
            self.nbt += nbt
        for p in params:
            if p.grad is None or not p.is_contiguous() or not p.grad.is_contiguous():
                raise _lib.FlowStateError("flowstate_b200: every conditioner / spline parameter needs a contiguous "
                                          ".grad tensor before the native training step is built")
        self._keep = (params, [p.grad for p in params], running)
        arr = lambda ts: (C.c_void_p * max(1, len(ts)))(*[t.data_ptr() for t in ts])
        ints = lambda t: (C.c_int * len(t))(*[int(v) for v in t.tolist()])
        self._arrays = (arr(params), arr([p.grad for p in params]), arr(running), ints(c0.transform_features),
                        ints(c0.identity_features))
        bn0 = net0.blocks[0].batch_norm_layers[0] if self.n_blocks else None
        d = _lib.FsTrainDesc()
        d.K, d.N, d.H, d.n_blocks, d.nb = self.K, self.N, self.H, self.n_blocks, self.nb
        d.bound = float(c0.tail_bound)
        d.feature_scale = float(net0.preprocessing.scale)
        d.bn_eps = float(bn0.eps) if bn0 is not None else 1e-3
        d.bn_momentum = float(bn0.momentum) if bn0 is not None else 0.1
        d.params = C.cast(self._arrays[0], C.POINTER(C.c_void_p))
        d.grads = C.cast(self._arrays[1], C.POINTER(C.c_void_p))
        d.bn_running = C.cast(self._arrays[2], C.POINTER(C.c_void_p))
        d.transform_features = C.cast(self._arrays[3], C.POINTER(C.c_int))
        d.identity_features = C.cast(self._arrays[4], C.POINTER(C.c_int))
        self.device = params[0].device
        _lib.bind_device(self.device)
        h = C.c_void_p()
        _lib.check(_lib.lib().fs_train_create(C.byref(d), C.byref(h)))
        self._h = h
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().fs_train_destroy(h)
            except Exception:
                pass
            self._h = None

    def pointers_valid(self):
        """The engine addresses parameters, gradients and buffers by pointer: any re-allocation invalidates it."""
        params, grads, running = self._keep
        return all(p.grad is g and p.data_ptr() == a for p, g, a in zip(params, grads, self._arrays[0])) and \
            all(r.data_ptr() == a for r, a in zip(running, self._arrays[2]))

    def step(self, batch, update_running=True):
        """Fills the gradient tensors, returns the loss (a 1-element device tensor owned by this object)."""
        batch = _lib.require_cuda(batch.to(torch.float32), "batch")
        if batch.dim() != 2 or batch.shape[1] != 2 * self.N:
            raise ValueError("Expected features = {}, got {}.".format(2 * self.N, tuple(batch.shape)))
        _lib.check(_lib.lib().fs_train_forward_kld(self._h, _lib.ptr(batch), int(batch.shape[0]), _lib.ptr(self.loss),
                                                   1 if update_running else 0, _lib.stream_ptr(self.device)))
        if update_running and self.nbt:
            torch._foreach_add_(self.nbt, 1)
        return self.loss
