"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the reference flow.

One coupling layer = the reference's CircularCoupledRationalQuadraticSpline,
restated functionally over a plain state_dict (reference key names, SURVEY.md
A.5), following (paths relative to <ref>/NF/normflows):

  flows/neural_spline/wrapper.py:98-275   layer: forward = prqct.inverse, inverse = prqct.forward
  flows/neural_spline/coupling.py:71-134  split / conditioner / splines / scatter / roll by D/2
  flows/neural_spline/coupling.py:156-170,335-368  params (B,N,3nb+1); widths,heights /= sqrt(H)
  flows/neural_spline/coupling.py:176-265 unconditional per-feature spline on the identity half
  utils/splines.py:16-222                 unconstrained + rational-quadratic spline (fwd / inverse)
  utils/nn.py:120-137                     features cat[cos(s x), sin(s x)], s = pi/bound
  nets/resnet.py:7-104                    Linear -> n x [BN,ReLU,Linear,BN,ReLU,Linear]+skip -> Linear
  Energy/Uniform.py:50-74                 base log-prob  -D log(2 bound)  or -inf
  core.py:178-214                         sample / log_prob over K layers

Fork quirks reproduced: same mask in every layer (odd indices
transformed), roll by D/2 after the density-direction layer / before the
sampling-direction layer, 3nb+1 parameters per coordinate with independent
boundary derivatives, last knot + 1e-6 in the bin search, |discriminant|,
BatchNorm eps 1e-3 in eval mode, unconditional spline not scaled by sqrt(H).

dtype=torch.float32 mirrors the reference's arithmetic; dtype=torch.float64 is
the "truth" used to attribute error.  Pinned against the reference's own
outputs in tests/golden/flow_*.npz.
"""
import math

import torch
import torch.nn.functional as F

MIN_W = 1e-3   # utils/splines.py:6-8
MIN_H = 1e-3
MIN_D = 1e-3


def _knots(unnorm, bound, min_size):
    """splines.py:117-127 (and 131-143): softmax -> floor -> cumsum -> affine to
    [-bound, bound] with forced end knots; sizes recomputed from the knots."""
    nb = unnorm.shape[-1]
    w = F.softmax(unnorm, dim=-1)
    w = min_size + (1 - min_size * nb) * w
    cum = torch.cumsum(w, dim=-1)
    cum = F.pad(cum, pad=(1, 0), mode="constant", value=0.0)
    cum = (bound - (-bound)) * cum + (-bound)
    cum[..., 0] = -bound
    cum[..., -1] = bound
    return cum, cum[..., 1:] - cum[..., :-1]


def rqs(x, uw, uh, ud, bound, inverse):
    """splines.py:91-222 on elements already known to be inside [-bound, bound].
    x (...,), uw/uh (..., nb), ud (..., nb+1).  Returns (y, logabsdet)."""
    cumw, widths = _knots(uw, bound, MIN_W)
    cumh, heights = _knots(uh, bound, MIN_H)
    derivs = MIN_D + F.softplus(ud)
    knots = (cumh if inverse else cumw).clone()
    knots[..., -1] += 1e-6                                   # splines.py:11-13
    k = (torch.sum(x[..., None] >= knots, dim=-1) - 1)[..., None]
    g = lambda t: t.gather(-1, k)[..., 0]
    x_k, w_k = g(cumw), g(widths)
    y_k, h_k = g(cumh), g(heights)
    s_k = g(heights / widths)
    d_k = g(derivs)
    d_k1 = g(derivs[..., 1:])
    if inverse:
        dy = x - y_k
        t = d_k + d_k1 - 2 * s_k
        a = dy * t + h_k * (s_k - d_k)
        b = h_k * d_k - dy * t
        c = -s_k * dy
        disc = (b.pow(2) - 4 * a * c).abs()                  # splines.py:171
        root = (2 * c) / (-b - torch.sqrt(disc))
        y = root * w_k + x_k
        tt = root * (1 - root)
        den = s_k + t * tt
        num = s_k.pow(2) * (d_k1 * root.pow(2) + 2 * s_k * tt + d_k * (1 - root).pow(2))
        return y, -(torch.log(num) - 2 * torch.log(den))
    th = (x - x_k) / w_k
    tt = th * (1 - th)
    num = h_k * (s_k * th.pow(2) + d_k * tt)
    den = s_k + (d_k + d_k1 - 2 * s_k) * tt
    y = y_k + num / den
    dnum = s_k.pow(2) * (d_k1 * th.pow(2) + 2 * s_k * tt + d_k * (1 - th).pow(2))
    return y, torch.log(dnum) - 2 * torch.log(den)


def unconstrained_rqs(x, uw, uh, ud, bound, inverse):
    """splines.py:16-88 for list-valued 'circular' tails: elements outside
    [-bound, bound] pass through with log-det 0; ud already has nb+1 entries
    and the padded copy made at :36-37 is never read (A.4-Q4)."""
    inside = (x >= -bound) & (x <= bound)
    y = x.clone()
    ld = torch.zeros_like(x)
    if inside.any():
        yi, ldi = rqs(x[inside], uw[inside, :], uh[inside, :], ud[inside, :], bound, inverse)
        y[inside] = yi
        ld[inside] = ldi
    return y, ld


class FlowSpec:
    """Shapes read off a reference state_dict."""

    def __init__(self, sd, bound):
        self.K = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("flows."))
        p = "flows.0.prqct."
        self.idf = sd[p + "identity_features"].long()
        self.trf = sd[p + "transform_features"].long()
        self.D = len(self.idf) + len(self.trf)
        self.H = sd[p + "transform_net.initial_layer.weight"].shape[0]
        self.nb = sd[p + "unconditional_transform.unnormalized_widths"].shape[1]
        self.n_blocks = 1 + max(int(k.split(".")[5]) for k in sd
                                if k.startswith(p + "transform_net.blocks."))
        self.bound = float(bound)


def conditioner(sd, p, ident, bound, dtype):
    """nn.py:120-137 + resnet.py:92-104 in eval mode (BN uses running stats)."""
    g = lambda k: sd[p + k].to(dtype)
    s = math.pi / bound
    feat = torch.cat([torch.cos(s * ident), torch.sin(s * ident)], dim=-1)
    h = F.linear(feat, g("initial_layer.weight"), g("initial_layer.bias"))
    b = 0
    while (p + "blocks.%d.linear_layers.0.weight" % b) in sd:
        q = "blocks.%d." % b
        t = h
        for j in (0, 1):
            t = F.batch_norm(t, g(q + "batch_norm_layers.%d.running_mean" % j),
                             g(q + "batch_norm_layers.%d.running_var" % j),
                             g(q + "batch_norm_layers.%d.weight" % j),
                             g(q + "batch_norm_layers.%d.bias" % j),
                             training=False, eps=1e-3)
            t = F.relu(t)
            t = F.linear(t, g(q + "linear_layers.%d.weight" % j), g(q + "linear_layers.%d.bias" % j))
        h = h + t
        b += 1
    return F.linear(h, g("final_layer.weight"), g("final_layer.bias"))


def _cond_spline(sd, p, spec, ident, tr, inverse, dtype):
    theta = conditioner(sd, p + "transform_net.", ident, spec.bound, dtype)
    theta = theta.reshape(tr.shape[0], tr.shape[1], -1)       # coupling.py:166
    nb = spec.nb
    rs = math.sqrt(spec.H)
    uw = theta[..., :nb] / rs                                  # coupling.py:340-342
    uh = theta[..., nb:2 * nb] / rs
    ud = theta[..., 2 * nb:]
    return unconstrained_rqs(tr, uw, uh, ud, spec.bound, inverse)


def _uncond_spline(sd, p, spec, ident, inverse, dtype):
    q = p + "unconditional_transform."
    B = ident.shape[0]
    e = lambda k: sd[q + k].to(dtype)[None].expand(B, -1, -1)  # coupling.py:208-238
    return unconstrained_rqs(ident, e("unnormalized_widths"), e("unnormalized_heights"),
                             e("unnormalized_derivatives"), spec.bound, inverse)


def layer_inverse(sd, i, spec, x, dtype=torch.float32):
    """Density direction: wrapper.inverse -> Coupling.forward (coupling.py:71-102)."""
    p = "flows.%d.prqct." % i
    ident = x[:, spec.idf]
    tr = x[:, spec.trf]
    tr2, ld = _cond_spline(sd, p, spec, ident, tr, False, dtype)
    ld = ld.sum(dim=1)
    id2, ld_id = _uncond_spline(sd, p, spec, ident, False, dtype)
    ld = ld + ld_id.sum(dim=1)
    out = torch.empty_like(x)
    out[:, spec.idf] = id2
    out[:, spec.trf] = tr2
    h = spec.D // 2
    out = torch.cat([out[:, h:], out[:, :h]], dim=1)           # coupling.py:100-101
    return out, ld


def layer_forward(sd, i, spec, z, dtype=torch.float32):
    """Sampling direction: wrapper.forward -> Coupling.inverse (coupling.py:104-134)."""
    p = "flows.%d.prqct." % i
    h = spec.D // 2
    z = torch.cat([z[:, h:], z[:, :h]], dim=1)                 # coupling.py:113-114
    ident = z[:, spec.idf]
    tr = z[:, spec.trf]
    id2, ld_id = _uncond_spline(sd, p, spec, ident, True, dtype)
    tr2, ld = _cond_spline(sd, p, spec, id2, tr, True, dtype)
    out = torch.empty_like(z)
    out[:, spec.idf] = id2
    out[:, spec.trf] = tr2
    return out, ld_id.sum(dim=1) + ld.sum(dim=1)


def base_log_prob(z, spec):
    """Energy/Uniform.py:50-74."""
    inb = ((z >= -spec.bound) & (z <= spec.bound)).all(dim=1)
    c = -spec.D * torch.log(torch.tensor(2 * spec.bound))
    lp = torch.full((z.shape[0],), float(c), dtype=z.dtype)
    lp[~inb] = -float("inf")
    return lp


def inverse_and_log_det(sd, spec, x, dtype=torch.float32):
    """core.py:71-86: layers K-1 .. 0 in the density direction."""
    z = x.to(dtype)
    ld = torch.zeros(len(z), dtype=dtype)
    for i in range(spec.K - 1, -1, -1):
        z, l = layer_inverse(sd, i, spec, z, dtype)
        ld = ld + l
    return z, ld


def log_prob(sd, spec, x, dtype=torch.float32):
    """core.py:198-214."""
    z, ld = inverse_and_log_det(sd, spec, x, dtype)
    return ld + base_log_prob(z, spec).to(dtype)


def forward_and_log_det(sd, spec, z, dtype=torch.float32):
    """core.py:28-56 / 178-196: layers 0 .. K-1 in the sampling direction."""
    x = z.to(dtype)
    ld = torch.zeros(len(x), dtype=dtype)
    for i in range(spec.K):
        x, l = layer_forward(sd, i, spec, x, dtype)
        ld = ld + l
    return x, ld
