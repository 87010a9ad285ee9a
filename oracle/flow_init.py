"""Oracle (TEST INFRASTRUCTURE): synthetic state_dict of a reference flow, without building any module.

Key names and shapes follow the reference's state_dict (prefix flows.<i>.prqct.), restricted to the
entries the oracle's functional flow (oracle/flow_ref.py) reads.  Values follow the reference's initialisation
(nets/resnet.py:33-35: second linear of a block U(-1e-3, 1e-3); flows/neural_spline/wrapper.py:181-185: final layer
weight 0, bias log(exp(1 - 1e-3) - 1); nn.Linear default U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BatchNorm weight 1 bias 0)
plus the N(0, sigma^2) perturbation and randomised BatchNorm statistics of the benchmark flows.
Used by `bench.py --impl reference` so that the CPU arm does not import the product package.
"""
import math

import torch


def synthetic_state_dict(n, K, blocks, H, nb, sigma, seed=0):
    g = torch.Generator().manual_seed(seed)
    D = 2 * n
    P = 3 * nb + 1

    def uni(shape, a):
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    def noisy(t):
        return t + sigma * torch.randn(t.shape, generator=g)

    sd = {}
    for i in range(K):
        p = "flows.%d.prqct." % i
        sd[p + "identity_features"] = torch.arange(0, D, 2)
        sd[p + "transform_features"] = torch.arange(1, D, 2)
        t = p + "transform_net."
        sd[t + "initial_layer.weight"] = noisy(uni((H, 2 * n), 1 / math.sqrt(2 * n)))
        sd[t + "initial_layer.bias"] = noisy(uni((H,), 1 / math.sqrt(2 * n)))
        for b in range(blocks):
            q = t + "blocks.%d." % b
            for j in (0, 1):
                sd[q + "batch_norm_layers.%d.weight" % j] = noisy(torch.ones(H))
                sd[q + "batch_norm_layers.%d.bias" % j] = noisy(torch.zeros(H))
                sd[q + "batch_norm_layers.%d.running_mean" % j] = 0.1 * torch.randn(H, generator=g)
                sd[q + "batch_norm_layers.%d.running_var" % j] = 0.5 + torch.rand(H, generator=g)
                a = 1 / math.sqrt(H) if j == 0 else 1e-3
                sd[q + "linear_layers.%d.weight" % j] = noisy(uni((H, H), a))
                sd[q + "linear_layers.%d.bias" % j] = noisy(uni((H,), a))
        sd[t + "final_layer.weight"] = noisy(torch.zeros(n * P, H))
        sd[t + "final_layer.bias"] = noisy(torch.full((n * P,), math.log(math.exp(1 - 1e-3) - 1)))
        u = p + "unconditional_transform."
        sd[u + "unnormalized_widths"] = noisy(torch.zeros(n, nb))
        sd[u + "unnormalized_heights"] = noisy(torch.zeros(n, nb))
        sd[u + "unnormalized_derivatives"] = noisy(torch.full((n, nb + 1), math.log(math.exp(1 - 1e-3) - 1)))
    return sd
