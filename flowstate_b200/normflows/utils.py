"""Helpers of NF/normflows/utils/{nn,masks}.py that the spline coupling uses."""
import torch
from torch import nn


def create_alternating_binary_mask(features, even=True):
    """utils/masks.py:4-17"""
    mask = torch.zeros(features).byte()
    mask[(0 if even else 1)::2] += 1
    return mask


def sum_except_batch(x, num_batch_dims=1):
    """utils/nn.py:197-200"""
    return torch.sum(x, dim=list(range(num_batch_dims, x.ndimension())))


class PeriodicFeaturesElementwise(nn.Module):
    """utils/nn.py:65-137 as modified by the fork: outputs cat[cos(s x), sin(s x)] over all
    inputs; `weights` is registered (and saved) but unused."""

    def __init__(self, ndim, ind, scale=1.0, bias=False, activation=None):
        super().__init__()
        self.ndim = ndim
        ind = torch.as_tensor(ind, dtype=torch.long)
        self.register_buffer("ind", ind)
        rest = [i for i in range(ndim) if i not in set(ind.tolist())]
        self.register_buffer("ind_", torch.tensor(rest, dtype=torch.long))
        perm = torch.cat((self.ind, self.ind_))
        inv = torch.zeros_like(perm)
        inv[perm] = torch.arange(ndim)
        self.register_buffer("inv_perm", inv)
        self.weights = nn.Parameter(torch.ones(len(self.ind), 2))
        if torch.is_tensor(scale):
            self.register_buffer("scale", scale)
        else:
            self.scale = scale
        self.apply_bias = bias
        if bias:
            self.bias = nn.Parameter(torch.zeros(len(self.ind)))
        self.activation = activation if activation is not None else nn.Identity()

    def forward(self, inputs):
        a = self.scale * inputs
        out = torch.cat([torch.cos(a), torch.sin(a)], dim=-1)
        if self.apply_bias:
            out = out + self.bias
        return self.activation(out)
