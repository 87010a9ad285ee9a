"""FlowPack: hands the parameters of K coupling layers to fs_flow_create and runs
fs_flow_inverse / fs_flow_forward on CUDA tensors (eval-mode inference only).

The pack is a snapshot: it records the autograd version counter of sentinel
tensors of every layer and is rebuilt when any of them changed (optimizer steps
and load_state_dict modify tensors in place, which bumps the counter); train(),
load_state_dict() and .to() drop it outright, NormalizingFlow.repack() does so
on request.
"""
import ctypes as C

import numpy as np
import os

import torch

from ._bridge import _lib

_PREC = {"fp32": _lib.FS_PREC_FP32, "tf32": _lib.FS_PREC_TF32}


def _layer_tensors(layer):
    c = layer.prqct
    net = c.transform_net
    ts = [net.initial_layer.weight, net.initial_layer.bias, net.final_layer.weight, net.final_layer.bias]
    for blk in net.blocks:
        for j in (0, 1):
            bn, lin = blk.batch_norm_layers[j], blk.linear_layers[j]
            ts += [bn.weight, bn.bias, bn.running_mean, bn.running_var, lin.weight, lin.bias]
    u = c.unconditional_transform
    ts += [u.unnormalized_widths, u.unnormalized_heights, u.unnormalized_derivatives]
    return ts


class FlowPack:
    def __init__(self, layers):
        if not layers:
            raise _lib.FlowStateError("flowstate_b200: cannot pack an empty flow")
        l0 = layers[0].prqct
        dev = l0.transform_net.initial_layer.weight.device
        if dev.type != "cuda":
            raise _lib.FlowStateError(
                "flowstate_b200: eval-mode flows run in CUDA kernels; move the model to a CUDA device "
                "(no CPU fallback)")
        self.device = dev
        self.K = len(layers)
        self.D = l0.features
        self.N = len(l0.transform_features)
        self.H = l0.transform_net.hidden_features
        self.nb = l0.num_bins
        self.n_blocks = len(l0.transform_net.blocks)
        self.bound = float(l0.tail_bound)
        self.precision = "auto"      # "auto": tensor cores when the flow shape has them | "fp32" | "tf32"
        self._sig = self._signature(layers)
        self._ws = {}
        self.launches = 0

        keep = []   # host arrays must outlive fs_flow_create

        def host(t):
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        def host_stack(ts):
            a = np.ascontiguousarray(np.stack([t.detach().to("cpu", torch.float32).numpy() for t in ts]))
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        arr = (_lib.FsLayerParams * self.K)()
        for i, layer in enumerate(layers):
            c = layer.prqct
            net = c.transform_net
            if (c.features != self.D or net.hidden_features != self.H or c.num_bins != self.nb
                    or len(net.blocks) != self.n_blocks or float(c.tail_bound) != self.bound
                    or not torch.equal(c.identity_features.cpu(), l0.identity_features.cpu())):
                raise _lib.FlowStateError("flowstate_b200: all coupling layers of a flow must share one shape")
            p = arr[i]
            p.init_w, p.init_b = host(net.initial_layer.weight), host(net.initial_layer.bias)
            if self.n_blocks:
                bns = [blk.batch_norm_layers[j] for blk in net.blocks for j in (0, 1)]
                lins = [blk.linear_layers[j] for blk in net.blocks for j in (0, 1)]
                p.bn_w = host_stack([b.weight for b in bns])
                p.bn_b = host_stack([b.bias for b in bns])
                p.bn_mean = host_stack([b.running_mean for b in bns])
                p.bn_var = host_stack([b.running_var for b in bns])
                p.lin_w = host_stack([l.weight for l in lins])
                p.lin_b = host_stack([l.bias for l in lins])
            p.final_w, p.final_b = host(net.final_layer.weight), host(net.final_layer.bias)
            u = c.unconditional_transform
            p.un_w, p.un_h, p.un_d = (host(u.unnormalized_widths), host(u.unnormalized_heights),
                                      host(u.unnormalized_derivatives))
        idf = np.ascontiguousarray(l0.identity_features.cpu().numpy().astype(np.int32))
        trf = np.ascontiguousarray(l0.transform_features.cpu().numpy().astype(np.int32))
        d = _lib.FsFlowDesc()
        d.K, d.N, d.H, d.n_blocks, d.nb = self.K, self.N, self.H, self.n_blocks, self.nb
        d.bound = self.bound
        d.bn_eps = float(l0.transform_net.blocks[0].batch_norm_layers[0].eps) if self.n_blocks else 1e-3
        d.identity_features = idf.ctypes.data_as(C.c_void_p)
        d.transform_features = trf.ctypes.data_as(C.c_void_p)
        d.layers = C.cast(arr, C.POINTER(_lib.FsLayerParams))
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.bind_device(dev)
            _lib.check(_lib.lib().fs_flow_create(C.byref(d), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().fs_flow_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- device-side refresh after the parameters changed (fs_flow_update) -----
    def _layer_sources(self, layer):
        """Source tensors of one layer in the order of the flat staging buffer (fs_layer_params order, the per-block
        tensors in [block, j] order so each stack is contiguous)."""
        c = layer.prqct
        net = c.transform_net
        bns = [blk.batch_norm_layers[j] for blk in net.blocks for j in (0, 1)]
        lins = [blk.linear_layers[j] for blk in net.blocks for j in (0, 1)]
        u = c.unconditional_transform
        groups = [[net.initial_layer.weight], [net.initial_layer.bias], [b.weight for b in bns], [b.bias for b in bns],
                  [b.running_mean for b in bns], [b.running_var for b in bns], [l.weight for l in lins],
                  [l.bias for l in lins], [net.final_layer.weight], [net.final_layer.bias],
                  [u.unnormalized_widths], [u.unnormalized_heights], [u.unnormalized_derivatives]]
        return groups

    def update(self, layers):
        """Refreshes the packed weights from the layers' current parameters without leaving the device: one
        multi-tensor copy into a flat staging buffer (the stacks fs_layer_params expects), then fs_flow_update."""
        names = ("init_w", "init_b", "bn_w", "bn_b", "bn_mean", "bn_var", "lin_w", "lin_b", "final_w", "final_b",
                 "un_w", "un_h", "un_d")
        srcs, sizes = [], []
        for layer in layers:
            groups = self._layer_sources(layer)
            sizes.append([[t.numel() for t in g] for g in groups])
            srcs.extend(t.detach() for g in groups for t in g)
        total = sum(n for lay in sizes for g in lay for n in g)
        if getattr(self, "_flat", None) is None or self._flat.numel() != total:
            self._flat = torch.empty(total, dtype=torch.float32, device=self.device)
            views, offs, off = [], [], 0
            for lay in sizes:
                lay_offs = []
                for g in lay:
                    lay_offs.append(off)
                    for n in g:
                        views.append(self._flat[off:off + n])
                        off += n
                offs.append(lay_offs)
            self._flat_views, self._flat_offs = views, offs
        torch._foreach_copy_(self._flat_views, [t.reshape(-1).float() if t.dtype != torch.float32 else t.reshape(-1)
                                                for t in srcs])
        arr = (_lib.FsLayerParams * self.K)()
        base = self._flat.data_ptr()
        for i in range(self.K):
            for name, off in zip(names, self._flat_offs[i]):
                setattr(arr[i], name, base + 4 * off)
        d = _lib.FsFlowDesc()
        d.K, d.N, d.H, d.n_blocks, d.nb = self.K, self.N, self.H, self.n_blocks, self.nb
        d.bound = self.bound
        l0 = layers[0].prqct
        d.bn_eps = float(l0.transform_net.blocks[0].batch_norm_layers[0].eps) if self.n_blocks else 1e-3
        d.layers = C.cast(arr, C.POINTER(_lib.FsLayerParams))
        _lib.check(_lib.lib().fs_flow_update(self._h, C.byref(d), _lib.stream_ptr(self.device)))
        self._sig = self._signature(layers)
        self.updates = getattr(self, "updates", 0) + 1

    def same_shape(self, layers):
        if len(layers) != self.K:
            return False
        for l in layers:
            c = l.prqct
            net = c.transform_net
            if (c.features != self.D or net.hidden_features != self.H or c.num_bins != self.nb
                    or len(net.blocks) != self.n_blocks or float(c.tail_bound) != self.bound
                    or net.initial_layer.weight.device != self.device):
                return False
        return True

    @staticmethod
    def _signature(layers):
        # a few sentinel tensors per layer: optimizer steps and load_state_dict touch all of them
        sig = []
        for l in layers:
            net = l.prqct.transform_net
            ts = [net.initial_layer.weight, net.final_layer.weight, net.final_layer.bias,
                  l.prqct.unconditional_transform.unnormalized_widths]
            if len(net.blocks):
                ts.append(net.blocks[0].batch_norm_layers[0].running_mean)
            sig.extend((id(t), t._version, t.data_ptr()) for t in ts)
        return tuple(sig)

    def matches(self, layers):
        return len(layers) == self.K and self._signature(layers) == self._sig

    def resolved_precision(self):
        if self.precision != "auto":
            return self.precision
        return "tf32" if _lib.lib().fs_flow_has_tensor_path(self._h) else "fp32"

    # -- calls ------------------------------------------------------------
    def _workspace(self, B, prec):
        # one buffer per (batch, precision, path, stream): passes issued on different streams may overlap
        key = (B, prec, bool(os.environ.get("FS_NO_FUSE")), bool(os.environ.get("FS_NO_LP")),
               torch.cuda.current_stream(self.device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None:
            n = _lib.lib().fs_flow_workspace_bytes(self._h, B, prec)
            ws = torch.empty(max(int(n), 16), dtype=torch.uint8, device=self.device)
            if len(self._ws) > 6:
                self._ws.clear()
            self._ws[key] = ws
        return ws

    def _prep(self, x):
        x = _lib.require_cuda(x.detach() if x.requires_grad else x, "flow input")
        if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.D:
            raise ValueError("Expected a float32 tensor of shape (B, %d), got %s %s" % (self.D, x.dtype, tuple(x.shape)))
        return x

    def inverse(self, x, want_logq=False, in_shift=0.0):
        """Density direction over all K layers: returns (z, logdet, logq or None)."""
        x = self._prep(x)
        B = x.shape[0]
        prec = _PREC[self.resolved_precision()]
        z = torch.empty_like(x)
        ld = torch.empty(B, dtype=torch.float32, device=x.device)
        lq = torch.empty(B, dtype=torch.float32, device=x.device) if want_logq else None
        nan = torch.zeros(1, dtype=torch.int32, device=x.device)
        if B:
            ws = self._workspace(B, prec)
            _lib.check(_lib.lib().fs_flow_inverse(self._h, _lib.ptr(x), B, float(in_shift), _lib.ptr(z), _lib.ptr(ld),
                                                  _lib.ptr(lq), _lib.ptr(nan), _lib.ptr(ws), ws.numel(), prec,
                                                  _lib.stream_ptr(x.device)))
        self._nan = nan
        return z, ld, lq

    def forward(self, z, want_logdet=True, out_shift=0.0):
        """Sampling direction over all K layers: returns (x, logdet or None)."""
        z = self._prep(z)
        B = z.shape[0]
        prec = _PREC[self.resolved_precision()]
        x = torch.empty_like(z)
        ld = torch.empty(B, dtype=torch.float32, device=z.device) if want_logdet else None
        nan = torch.zeros(1, dtype=torch.int32, device=z.device)
        if B:
            ws = self._workspace(B, prec)
            _lib.check(_lib.lib().fs_flow_forward(self._h, _lib.ptr(z), B, float(out_shift), _lib.ptr(x), _lib.ptr(ld),
                                                  _lib.ptr(nan), _lib.ptr(ws), ws.numel(), prec,
                                                  _lib.stream_ptr(z.device)))
        self._nan = nan
        return x, ld

    def conditioner(self, layer, features):
        """Conditioner (ResidualNet, eval mode) of one layer on periodic features [rows, 2N] -> theta."""
        features = _lib.require_cuda(features, "features")
        rows = features.shape[0]
        prec = _PREC[self.resolved_precision()]
        theta = torch.empty(rows, self.N * (3 * self.nb + 1), dtype=torch.float32, device=features.device)
        ws = self._workspace(rows, prec)
        _lib.check(_lib.lib().fs_flow_conditioner(self._h, int(layer), _lib.ptr(features), rows, _lib.ptr(theta),
                                                  _lib.ptr(ws), ws.numel(), prec, _lib.stream_ptr(features.device)))
        return theta

    def tile_features(self, features):
        """Row-major periodic features [rows, 2N] -> the row-tiled layout the tensor path reads (fs_flow_tile_features);
        pass the result to coupling(..., tiled=True)."""
        features = _lib.require_cuda(features, "features")
        rows, k0 = features.shape
        nbytes = _lib.lib().fs_flow_tiled_features_bytes(int(rows), int(k0))
        out = torch.zeros(nbytes // 4, dtype=torch.float32, device=features.device)
        _lib.check(_lib.lib().fs_flow_tile_features(_lib.ptr(features), int(rows), int(k0), _lib.ptr(out),
                                                    _lib.stream_ptr(features.device)))
        return out

    def coupling(self, layer, direction, features, xin, xout=None, logdet=None, tiled=False):
        """Conditioner + conditional spline of the transformed half of one layer in one kernel (fused
        tensor-core path; coupling.py:86-102 density / 126-135 sampling).  Returns (xout, logdet).
        tiled: `features` is the output of tile_features (the layout the full passes use) instead of [rows, 2N]."""
        features = _lib.require_cuda(features, "features")
        xin = _lib.require_cuda(xin, "xin")
        rows = xin.shape[0]
        if xout is None:
            xout = torch.zeros_like(xin)
        if logdet is None:
            logdet = torch.zeros(rows, dtype=torch.float32, device=xin.device)
        code = {"density": 1, "sampling": 2}[direction] | (0x10 if tiled else 0)
        _lib.check(_lib.lib().fs_flow_coupling(self._h, int(layer), code, _lib.ptr(features), _lib.ptr(xin),
                                               _lib.ptr(xout), _lib.ptr(logdet), rows, _lib.ptr(self._nan),
                                               _lib.stream_ptr(xin.device)))
        return xout, logdet

    def coupling_all(self, direction, features_all, buf0, buf1=None):
        """All K layers' conditioners + conditional splines in one launch (fs_flow_coupling_all).  features_all: flat
        tensor of K tiled feature matrices in step order (tile_features per layer, concatenated).  Returns
        (buf holding the result, logdet partials [K, rows])."""
        buf0 = _lib.require_cuda(buf0, "buf0")
        rows = buf0.shape[0]
        if buf1 is None:
            buf1 = torch.zeros_like(buf0)
        parts = torch.empty(self.K, rows, dtype=torch.float32, device=buf0.device)
        scratch = torch.empty(self.K * ((rows + 127) // 128) * 8, dtype=torch.int32, device=buf0.device)
        code = {"density": 1, "sampling": 2}[direction]
        _lib.check(_lib.lib().fs_flow_coupling_all(self._h, code, _lib.ptr(features_all), _lib.ptr(buf0), _lib.ptr(buf1),
                                                   _lib.ptr(parts), _lib.ptr(scratch), rows, _lib.ptr(self._nan),
                                                   _lib.stream_ptr(buf0.device)))
        return (buf1 if self.K & 1 else buf0), parts

    def set_layer_parallel(self, mode):
        """Scheduling hint (fs_flow_set_layer_parallel): "auto" | "prefer" (this flow's passes run alone on the GPU) |
        "never"."""
        code = {"auto": 0, "prefer": 1, "never": 2}[mode]
        if getattr(self, "_lp_mode", 0) != code:
            _lib.check(_lib.lib().fs_flow_set_layer_parallel(self._h, code))
            self._lp_mode = code

    def uses_layer_parallel(self, rows):
        """True when log_prob / sample passes of `rows` rows run as one layer-parallel launch (fs_flow_uses_layer_parallel)."""
        return bool(_lib.lib().fs_flow_uses_layer_parallel(self._h, int(rows), _PREC[self.resolved_precision()]))

    def check_nan(self):
        """Surfaces the device-side NaN flag like the reference's ValueError (utils/splines.py:176-183)."""
        if int(self._nan.item()):
            raise ValueError("Discriminant computation resulted in NaN.")
