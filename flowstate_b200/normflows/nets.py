"""Residual conditioner, module tree of NF/normflows/nets/resnet.py:7-104
(same attribute names -> same state_dict keys, same construction order -> same
random initialisation for a given torch seed)."""
import torch
from torch import nn
from torch.nn import functional as F, init


class ResidualBlock(nn.Module):
    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False, zero_initialization=True):
        super().__init__()
        if context_features is not None:
            raise NotImplementedError("context features are not used by the flow-state drivers")
        self.activation = activation
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(features, eps=1e-3) for _ in range(2)])
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        t = inputs
        for j in (0, 1):
            if self.use_batch_norm:
                t = self.batch_norm_layers[j](t)
            t = self.activation(t)
            if j == 1:
                t = self.dropout(t)
            t = self.linear_layers[j](t)
        return inputs + t


class ResidualNet(nn.Module):
    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False, preprocessing=None):
        super().__init__()
        if context_features is not None:
            raise NotImplementedError("context features are not used by the flow-state drivers")
        self.hidden_features = hidden_features
        self.context_features = context_features
        self.preprocessing = preprocessing
        self.initial_layer = nn.Linear(in_features, hidden_features)
        self.blocks = nn.ModuleList([
            ResidualBlock(hidden_features, context_features, activation=activation,
                          dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
            for _ in range(num_blocks)])
        self.final_layer = nn.Linear(hidden_features, out_features)

    def forward(self, inputs, context=None):
        t = inputs if self.preprocessing is None else self.preprocessing(inputs)
        t = self.initial_layer(t)
        for block in self.blocks:
            t = block(t)
        return self.final_layer(t)


class MLP(nn.Module):
    """nets/mlp.py:5-58: Linear / LeakyReLU stack, optional output function; the conditioner of the affine flows."""

    def __init__(self, layers, leaky=0.0, score_scale=None, output_fn=None, output_scale=None, init_zeros=False,
                 dropout=None):
        super().__init__()
        net = nn.ModuleList([])
        for k in range(len(layers) - 2):
            net.append(nn.Linear(layers[k], layers[k + 1]))
            net.append(nn.LeakyReLU(leaky))
        if dropout is not None:
            net.append(nn.Dropout(p=dropout))
        net.append(nn.Linear(layers[-2], layers[-1]))
        if init_zeros:
            nn.init.zeros_(net[-1].weight)
            nn.init.zeros_(net[-1].bias)
        if output_fn is not None:
            if score_scale is not None:
                net.append(_ConstScale(score_scale))
            if output_fn == "sigmoid":
                net.append(nn.Sigmoid())
            elif output_fn == "relu":
                net.append(nn.ReLU())
            elif output_fn == "tanh":
                net.append(nn.Tanh())
            elif output_fn == "clampexp":
                net.append(_ClampExp())
            if output_scale is not None:
                net.append(_ConstScale(output_scale))
        self.net = nn.Sequential(*net)

    def forward(self, x):
        return self.net(x)


class _ConstScale(nn.Module):
    """utils/nn.py ConstScaleLayer"""

    def __init__(self, scale=1.0):
        super().__init__()
        self.register_buffer("scale", torch.tensor(scale))

    def forward(self, x):
        return x * self.scale


class _ClampExp(nn.Module):
    """utils/nn.py ClampExp: exp(min(x, 0)), i.e. exp clamped to at most 1"""

    def forward(self, x):
        return torch.min(torch.exp(x), torch.tensor(1.0, device=x.device, dtype=x.dtype))
