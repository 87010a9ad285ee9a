"""Oracle (TEST INFRASTRUCTURE): numpy/torch-CPU restatement of the affine (RealNVP) coupling transforms and periodic
shifts of the reference.

  NF/normflows/flows/affine/coupling.py:163-229  MaskedAffineFlow   f(z) = b z + (1 - b) (z exp(s(b z)) + t(b z));
                                                 non-finite s / t -> NaN; log-det sum((1 - b) s)
  NF/normflows/flows/affine/coupling.py:99-160   AffineCoupling     shift = param[:, 0::2], scale = param[:, 1::2];
                                                 scale maps exp / sigmoid (z / sigmoid(s + 2) + t) / sigmoid_inv
  NF/normflows/flows/periodic.py:6-73            PeriodicWrap / PeriodicShift   remainder(z + shift + bound, 2 bound) - bound
  NF/normflows/nets/mlp.py:5-58                  MLP (Linear / LeakyReLU stack), evaluated from a state_dict
Pinned against outputs of the reference in tests/golden/affine.npz (oracle/make_golden.py: gen_affine).
"""
import torch
import torch.nn.functional as F


def mlp(sd, prefix, x, leaky=0.0):
    """nets/mlp.py: net.<2k> are the Linear layers, LeakyReLU(leaky) between them."""
    k = 0
    while (prefix + "net.%d.weight" % (k + 2)) in sd:
        x = F.leaky_relu(F.linear(x, sd[prefix + "net.%d.weight" % k].to(x.dtype), sd[prefix + "net.%d.bias" % k].to(x.dtype)), leaky)
        k += 2
    return F.linear(x, sd[prefix + "net.%d.weight" % k].to(x.dtype), sd[prefix + "net.%d.bias" % k].to(x.dtype))


def masked_affine(z, b, scale, trans, inverse):
    nan = torch.tensor(float("nan"), dtype=z.dtype)
    scale = torch.where(torch.isfinite(scale), scale, nan)
    trans = torch.where(torch.isfinite(trans), trans, nan)
    zm = b * z
    if inverse:
        return zm + (1 - b) * (z - trans) * torch.exp(-scale), -torch.sum((1 - b) * scale, dim=1)
    return zm + (1 - b) * (z * torch.exp(scale) + trans), torch.sum((1 - b) * scale, dim=1)


def affine_coupling(z2, param, scale_map, inverse):
    shift, s = param[:, 0::2], param[:, 1::2]
    if scale_map == "exp":
        out = (z2 - shift) * torch.exp(-s) if inverse else z2 * torch.exp(s) + shift
        ld = torch.sum(s, dim=1)
        return out, (-ld if inverse else ld)
    sg = torch.sigmoid(s + 2)
    mul = (scale_map == "sigmoid_inv") != inverse
    if inverse:
        out = (z2 - shift) * sg if mul else (z2 - shift) / sg
    else:
        out = z2 * sg + shift if mul else z2 / sg + shift
    ld = torch.sum(torch.log(sg), dim=1)
    return out, (ld if mul else -ld)


def periodic_shift(z, ind, bound, shift):
    out = z.clone()
    out[..., ind] = torch.remainder(out[..., ind] + shift + bound, 2 * bound) - bound
    return out
