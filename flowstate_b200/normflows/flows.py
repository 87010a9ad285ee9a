"""Circular coupled rational-quadratic-spline layer.

Module interface and parameter tree of the reference's
CircularCoupledRationalQuadraticSpline
(NF/normflows/flows/neural_spline/wrapper.py:98-275) and the coupling it wraps
(flows/neural_spline/coupling.py:16-368):

    flows.<i>.prqct.identity_features / transform_features          (buffers)
    flows.<i>.prqct.transform_net.{preprocessing, initial_layer, blocks.<b>.*, final_layer}
    flows.<i>.prqct.unconditional_transform.unnormalized_{widths,heights,derivatives}

forward(z) is the sampling direction, inverse(z) the density direction; both
return (tensor (B, D), log-det (B,)).  Eval mode + CUDA -> fs_flow_forward /
fs_flow_inverse on a one-layer pack; train mode -> autograd path.
"""
import math

import numpy as np
import torch
from torch import nn

from . import _spline_torch
from .nets import ResidualNet
from .utils import PeriodicFeaturesElementwise, create_alternating_binary_mask

DEFAULT_MIN_DERIVATIVE = 1e-3


class Flow(nn.Module):
    """flows/base.py:5-24"""

    def forward(self, z):
        raise NotImplementedError("Forward pass has not been implemented.")

    def inverse(self, z):
        raise NotImplementedError("This flow has no algebraic inverse.")


class _UnconditionalSpline(nn.Module):
    """Parameter holder of coupling.py:176-265 (identity initialisation, nb + 1 derivatives
    because `tails` is a list)."""

    def __init__(self, features, num_bins):
        super().__init__()
        self.unnormalized_widths = nn.Parameter(torch.zeros(features, num_bins))
        self.unnormalized_heights = nn.Parameter(torch.zeros(features, num_bins))
        c = np.log(np.exp(1 - DEFAULT_MIN_DERIVATIVE) - 1)
        self.unnormalized_derivatives = nn.Parameter(c * torch.ones(features, num_bins + 1))


class _SplineCoupling(nn.Module):
    """`prqct` of the reference layer."""

    def __init__(self, mask, net_fn, num_bins, tail_bound):
        super().__init__()
        mask = torch.as_tensor(mask)
        if mask.dim() != 1:
            raise ValueError("Mask must be a 1-dim tensor.")
        if mask.numel() <= 0:
            raise ValueError("Mask can't be empty.")
        self.features = len(mask)
        fv = torch.arange(self.features)
        self.register_buffer("identity_features", fv.masked_select(mask <= 0))
        self.register_buffer("transform_features", fv.masked_select(mask > 0))
        self.num_bins = num_bins
        self.tail_bound = tail_bound
        n_id, n_tr = len(self.identity_features), len(self.transform_features)
        self.transform_net = net_fn(n_id, n_tr * (3 * num_bins + 1))
        self.unconditional_transform = _UnconditionalSpline(n_id, num_bins)


class CircularCoupledRationalQuadraticSpline(Flow):
    def __init__(self, num_input_channels, num_blocks, num_hidden_channels, ind_circ, num_heads=4,
                 num_context_channels=None, num_bins=8, tail_bound=3.0, net_type="residual", activation=nn.ReLU,
                 dropout_probability=0.0, reverse_mask=False, mask=None, init_identity=True):
        super().__init__()
        if net_type not in ("residual", None):
            raise NotImplementedError("only the residual conditioner (the drivers' default) is built")
        if torch.is_tensor(tail_bound) or num_context_channels is not None:
            raise NotImplementedError("tensor tail bounds / context channels are not used by the flow-state drivers")
        if num_input_channels % 2:
            raise ValueError("num_input_channels must be even (2 coordinates per particle)")
        if mask is None:
            mask = create_alternating_binary_mask(num_input_channels, even=reverse_mask)
        identity = torch.arange(num_input_channels).masked_select(torch.as_tensor(mask) <= 0)
        circ = set(int(i) for i in ind_circ)
        if any(int(i) not in circ for i in range(num_input_channels)):
            raise NotImplementedError("every coordinate must be circular (as in the flow-state drivers)")
        ind_circ_id = list(range(len(identity)))
        scale_pf = np.pi / tail_bound

        def net_fn(in_features, out_features):
            pf = PeriodicFeaturesElementwise(in_features, ind_circ_id, scale_pf)
            net = ResidualNet(in_features=2 * in_features, out_features=out_features,
                              context_features=None, hidden_features=num_hidden_channels, num_blocks=num_blocks,
                              activation=activation(), dropout_probability=dropout_probability,
                              use_batch_norm=True, preprocessing=pf)
            if init_identity:
                nn.init.constant_(net.final_layer.weight, 0.0)
                nn.init.constant_(net.final_layer.bias, np.log(np.exp(1 - DEFAULT_MIN_DERIVATIVE) - 1))
            return net

        self.prqct = _SplineCoupling(mask, net_fn, num_bins, tail_bound)
        self._pack = None
        self.precision = "auto"      # conditioner arithmetic of the eval-mode kernels, as NormalizingFlow.precision
        self.fused_training = True   # train mode on CUDA: splines through fs_spline_train_fwd / _bwd (False: torch ops)

    # -- shapes -----------------------------------------------------------
    @property
    def bound(self):
        return float(self.prqct.tail_bound)

    def _check(self, z):
        if z.dim() != 2:
            raise ValueError("Inputs must be a 2D or a 4D tensor.")
        if z.shape[1] != self.prqct.features:
            raise ValueError("Expected features = {}, got {}.".format(self.prqct.features, z.shape[1]))

    # -- train-mode (autograd) path --------------------------------------
    def _params(self, ident):
        c = self.prqct
        theta = c.transform_net(ident)
        theta = theta.reshape(ident.shape[0], len(c.transform_features), -1)
        nb = c.num_bins
        rs = math.sqrt(c.transform_net.hidden_features)
        return theta[..., :nb] / rs, theta[..., nb:2 * nb] / rs, theta[..., 2 * nb:]

    def _uncond(self, ident, inverse):
        u = self.prqct.unconditional_transform
        B = ident.shape[0]
        e = lambda p: p[None].expand(B, *p.shape)
        return _spline_torch.spline(ident, e(u.unnormalized_widths), e(u.unnormalized_heights),
                                    e(u.unnormalized_derivatives), self.bound, inverse)

    def _density_torch(self, x):
        c = self.prqct
        ident, tr = x[:, c.identity_features], x[:, c.transform_features]
        if x.is_cuda and x.dtype == torch.float32 and self.fused_training:
            # training on the device: both splines through the hand-written forward / backward kernels
            nb = c.num_bins
            theta = c.transform_net(ident).reshape(x.shape[0], len(c.transform_features), 3 * nb + 1)
            tr2, ld = _spline_torch.fused_spline(tr, theta, self.bound, nb,
                                                 1.0 / math.sqrt(c.transform_net.hidden_features))
            u = c.unconditional_transform
            shared = torch.cat([u.unnormalized_widths, u.unnormalized_heights, u.unnormalized_derivatives], dim=1)
            id2, ld_id = _spline_torch.fused_spline(ident, shared, self.bound, nb, 1.0)
        else:
            uw, uh, ud = self._params(ident)
            tr2, ld = _spline_torch.spline(tr, uw, uh, ud, self.bound, False)
            id2, ld_id = self._uncond(ident, False)
        out = torch.empty_like(x)
        out[:, c.identity_features] = id2
        out[:, c.transform_features] = tr2
        h = c.features // 2
        return torch.cat([out[:, h:], out[:, :h]], dim=1), ld.sum(1) + ld_id.sum(1)

    def _sampling_torch(self, z):
        c = self.prqct
        h = c.features // 2
        z = torch.cat([z[:, h:], z[:, :h]], dim=1)
        ident, tr = z[:, c.identity_features], z[:, c.transform_features]
        id2, ld_id = self._uncond(ident, True)
        uw, uh, ud = self._params(id2)
        tr2, ld = _spline_torch.spline(tr, uw, uh, ud, self.bound, True)
        out = torch.empty_like(z)
        out[:, c.identity_features] = id2
        out[:, c.transform_features] = tr2
        return out, ld_id.sum(1) + ld.sum(1)

    # -- public interface -------------------------------------------------
    def _cuda_pack(self):
        from ._pack import FlowPack
        if self._pack is None or not self._pack.same_shape([self]):
            self._pack = FlowPack([self])
        elif not self._pack.matches([self]):
            self._pack.update([self])
        self._pack.precision = self.precision
        return self._pack

    def forward(self, z, context=None):
        self._check(z)
        if self.training:
            return self._sampling_torch(z)
        x, ld = self._cuda_pack().forward(z, want_logdet=True)
        return x, ld.view(-1)

    def inverse(self, z, context=None):
        self._check(z)
        if self.training:
            return self._density_torch(z)
        x, ld, _ = self._cuda_pack().inverse(z)
        return x, ld.view(-1)


# affine (RealNVP) coupling and periodic flows under the reference's names
from .affine import (AffineCoupling, AffineCouplingBlock, MaskedAffineFlow, Merge, PeriodicShift,  # noqa: E402,F401
                     PeriodicWrap, Split)
