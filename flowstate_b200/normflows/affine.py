"""Affine (RealNVP) coupling layers and periodic shifts / wraps.

Mirrors NF/normflows/flows/affine/coupling.py:99-268 (AffineCoupling, MaskedAffineFlow, AffineCouplingBlock),
flows/reshape.py (Split / Merge, channel modes) and flows/periodic.py:6-73 (PeriodicWrap, PeriodicShift) with the
reference's constructor signatures and module names.  The conditioner networks (`param_map`, `s`, `t`) stay ordinary
torch modules (plain library GEMMs); on CUDA float32 tensors without gradient the element-wise transform, the mask,
the non-finite -> NaN rule and the per-row log-determinant run in ONE kernel (fs_affine_coupling), the periodic maps in
fs_periodic_shift.  With autograd enabled (training) the torch expressions of the reference are used.
"""
import numpy as np
import torch
import torch.nn as nn

from ._bridge import _lib

_SCALE_MAPS = {"exp": 0, "sigmoid": 1, "sigmoid_inv": 2}


def _kernel_ok(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        return False
    return all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors) and tensors[0].dim() == 2


def _affine(z, mask, scale, shift, scale_map, inverse, nan_rule, es=1):
    """fs_affine_coupling on [rows, n] tensors; scale / shift are read with element stride `es`."""
    z = z.contiguous()
    rows, n = z.shape
    out = torch.empty_like(z)
    ld = torch.empty(rows, dtype=torch.float32, device=z.device)
    sc = scale.contiguous() if scale is not None else None
    sh = shift.contiguous() if shift is not None else None
    _lib.check(_lib.lib().fs_affine_coupling(
        _lib.ptr(z), n, _lib.ptr(mask), _lib.ptr(sc), sc.shape[1] if sc is not None else 0, es, _lib.ptr(sh),
        sh.shape[1] if sh is not None else 0, es, rows, n, _SCALE_MAPS[scale_map], int(inverse), int(nan_rule),
        _lib.ptr(out), n, _lib.ptr(ld), _lib.stream_ptr(z.device)))
    return out, ld


class Flow(nn.Module):
    """flows/base.py:5-24"""

    def forward(self, z):
        raise NotImplementedError("Forward pass has not been implemented.")

    def inverse(self, z):
        raise NotImplementedError("This flow has no algebraic inverse.")


class MaskedAffineFlow(Flow):
    """coupling.py:163-229: f(z) = b z + (1 - b) (z exp(s(b z)) + t(b z))."""

    def __init__(self, b, t=None, s=None):
        super().__init__()
        self.b_cpu = b.view(1, *b.size())
        self.register_buffer("b", self.b_cpu)
        if s is None:
            self.s = torch.zeros_like
        else:
            self.add_module("s", s)
        if t is None:
            self.t = torch.zeros_like
        else:
            self.add_module("t", t)

    def _apply_map(self, z, inverse):
        z_masked = self.b * z
        scale = self.s(z_masked)
        trans = self.t(z_masked)
        if _kernel_ok(z, scale, trans) and self.b.dim() == 2:
            return _affine(z, self.b.reshape(-1).float().contiguous(), scale, trans, "exp", inverse, True)
        nan = torch.tensor(np.nan, dtype=z.dtype, device=z.device)
        scale = torch.where(torch.isfinite(scale), scale, nan)
        trans = torch.where(torch.isfinite(trans), trans, nan)
        dims = list(range(1, self.b.dim()))
        if inverse:
            return (z_masked + (1 - self.b) * (z - trans) * torch.exp(-scale),
                    -torch.sum((1 - self.b) * scale, dim=dims))
        return z_masked + (1 - self.b) * (z * torch.exp(scale) + trans), torch.sum((1 - self.b) * scale, dim=dims)

    def forward(self, z):
        return self._apply_map(z, False)

    def inverse(self, z):
        return self._apply_map(z, True)


class AffineCoupling(Flow):
    """coupling.py:99-160: z = [z1, z2]; z1 conditions an affine map of z2 (shift = param[:, 0::2], scale = param[:, 1::2])."""

    def __init__(self, param_map, scale=True, scale_map="exp"):
        super().__init__()
        self.add_module("param_map", param_map)
        self.scale = scale
        self.scale_map = scale_map
        if scale and scale_map not in _SCALE_MAPS:
            raise NotImplementedError("This scale map is not implemented.")

    def _apply_map(self, z, inverse):
        z1, z2 = z
        param = self.param_map(z1)
        if _kernel_ok(z2, param):
            p = param.contiguous()
            if self.scale:        # interleaved parameters: element stride 2, shift at offset 0, scale at offset 1
                z2c = z2.contiguous()
                rows, n = z2c.shape
                out = torch.empty_like(z2c)
                ld = torch.empty(rows, dtype=torch.float32, device=z2c.device)
                base = p.data_ptr()
                import ctypes as C
                _lib.check(_lib.lib().fs_affine_coupling(
                    _lib.ptr(z2c), n, None, C.c_void_p(base + 4), p.shape[1], 2, C.c_void_p(base), p.shape[1], 2, rows, n,
                    _SCALE_MAPS[self.scale_map], int(inverse), 0, _lib.ptr(out), n, _lib.ptr(ld),
                    _lib.stream_ptr(z2c.device)))
                return [z1, out], ld
            out, ld = _affine(z2, None, None, p, "exp", inverse, False)
            return [z1, out], ld
        if self.scale:
            shift, scale_ = param[:, 0::2, ...], param[:, 1::2, ...]
            dims = list(range(1, shift.dim()))
            if self.scale_map == "exp":
                z2 = (z2 - shift) * torch.exp(-scale_) if inverse else z2 * torch.exp(scale_) + shift
                log_det = torch.sum(scale_, dim=dims)
                log_det = -log_det if inverse else log_det
            else:
                sg = torch.sigmoid(scale_ + 2)
                mul = (self.scale_map == "sigmoid_inv") != inverse
                if inverse:
                    z2 = (z2 - shift) * sg if mul else (z2 - shift) / sg
                else:
                    z2 = z2 * sg + shift if mul else z2 / sg + shift
                log_det = torch.sum(torch.log(sg), dim=dims)
                log_det = log_det if mul else -log_det
        else:
            z2 = z2 - param if inverse else z2 + param
            log_det = torch.zeros(len(z2), dtype=z2.dtype, device=z2.device)
        return [z1, z2], log_det

    def forward(self, z):
        return self._apply_map(z, False)

    def inverse(self, z):
        return self._apply_map(z, True)


class Split(Flow):
    """flows/reshape.py Split, channel modes (the checkerboard modes are for images)."""

    def __init__(self, mode="channel"):
        super().__init__()
        if mode not in ("channel", "channel_inv"):
            raise NotImplementedError("Mode " + mode + " is not implemented.")
        self.mode = mode

    def forward(self, z):
        a, b = z.chunk(2, dim=1)
        return ([a, b] if self.mode == "channel" else [b, a]), 0

    def inverse(self, z):
        z1, z2 = z
        return torch.cat([z1, z2] if self.mode == "channel" else [z2, z1], 1), 0


class Merge(Split):
    def forward(self, z):
        return super().inverse(z)

    def inverse(self, z):
        return super().forward(z)


class AffineCouplingBlock(Flow):
    """coupling.py:232-268: Split -> AffineCoupling -> Merge."""

    def __init__(self, param_map, scale=True, scale_map="exp", split_mode="channel"):
        super().__init__()
        self.flows = nn.ModuleList([Split(split_mode), AffineCoupling(param_map, scale, scale_map), Merge(split_mode)])

    def forward(self, z):
        log_det_tot = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for flow in self.flows:
            z, log_det = flow(z)
            log_det_tot = log_det_tot + log_det
        return z, log_det_tot

    def inverse(self, z):
        log_det_tot = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for i in range(len(self.flows) - 1, -1, -1):
            z, log_det = self.flows[i].inverse(z)
            log_det_tot = log_det_tot + log_det
        return z, log_det_tot


class _Periodic(Flow):
    def __init__(self, ind, bound):
        super().__init__()
        self.ind = ind
        if torch.is_tensor(bound):
            self.register_buffer("bound", bound)
        else:
            self.bound = bound
        self._tables = None

    def _shift_map(self, z, shift):
        """remainder(z[ind] + shift + bound, 2 bound) - bound on the selected columns."""
        if _kernel_ok(z):
            D = z.shape[1]
            ind = list(self.ind)
            key = (D, z.device, float(torch.as_tensor(shift).sum()))
            if self._tables is None or self._tables[0] != key:
                slot = torch.full((D,), -1, dtype=torch.int32)
                slot[torch.as_tensor(ind, dtype=torch.long)] = torch.arange(len(ind), dtype=torch.int32)
                b = torch.as_tensor(self.bound, dtype=torch.float32).reshape(-1)
                b = b.expand(len(ind)) if b.numel() == 1 else b
                sh = torch.as_tensor(shift, dtype=torch.float32).reshape(-1)
                sh = sh.expand(len(ind)) if sh.numel() == 1 else sh
                self._tables = (key, slot.to(z.device), b.contiguous().to(z.device), sh.contiguous().to(z.device))
            _, slot, b, sh = self._tables
            zc = z.contiguous()
            out = torch.empty_like(zc)
            _lib.check(_lib.lib().fs_periodic_shift(_lib.ptr(zc), zc.shape[0], D, _lib.ptr(slot), _lib.ptr(b),
                                                    _lib.ptr(sh), _lib.ptr(out), _lib.stream_ptr(z.device)))
            return out
        z_ = z.clone()
        z_[..., self.ind] = torch.remainder(z_[..., self.ind] + shift + self.bound, 2 * self.bound) - self.bound
        return z_


class PeriodicWrap(_Periodic):
    """periodic.py:6-33: forward is the identity, inverse wraps the selected coordinates into [-bound, bound)."""

    def __init__(self, ind, bound=1.0):
        super().__init__(ind, bound)

    def forward(self, z):
        return z, torch.zeros(len(z), dtype=z.dtype, device=z.device)

    def inverse(self, z):
        return self._shift_map(z, 0.0), torch.zeros(len(z), dtype=z.dtype, device=z.device)


class PeriodicShift(_Periodic):
    """periodic.py:36-73: shift and wrap; the inverse shifts back."""

    def __init__(self, ind, bound=1.0, shift=0.0):
        super().__init__(ind, bound)
        if torch.is_tensor(shift):
            self.register_buffer("shift", shift)
        else:
            self.shift = shift

    def forward(self, z):
        return self._shift_map(z, self.shift), torch.zeros(len(z), dtype=z.dtype, device=z.device)

    def inverse(self, z):
        sh = -self.shift
        return self._shift_map(z, sh), torch.zeros(len(z), dtype=z.dtype, device=z.device)
