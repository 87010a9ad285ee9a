"""Differentiable (autograd) rational-quadratic spline used ONLY in train mode.

Inference (eval mode) never comes here - it runs in the CUDA kernels.  The
arithmetic follows NF/normflows/utils/splines.py:16-222 with the fork's quirks
(last knot + 1e-6 in the bin search, |discriminant|, 3nb+1
parameters with independent boundary derivatives), written with torch.where
instead of boolean-mask gathers so it stays shape-static for training.
"""
import torch
import torch.nn.functional as F

MIN_SIZE = 1e-3
MIN_DERIV = 1e-3


def _cumulative(unnorm, bound):
    nb = unnorm.shape[-1]
    p = MIN_SIZE + (1.0 - MIN_SIZE * nb) * F.softmax(unnorm, dim=-1)
    c = F.pad(torch.cumsum(p, dim=-1), (1, 0))
    c = 2.0 * bound * c - bound
    edge = torch.zeros_like(c)
    edge[..., 0] = 1.0
    edge[..., -1] = 1.0
    target = torch.full_like(c, bound)
    target[..., 0] = -bound
    c = torch.where(edge.bool(), target, c)
    return c, c[..., 1:] - c[..., :-1]


def spline(x, uw, uh, ud, bound, inverse):
    """x (..., ), uw/uh (..., nb), ud (..., nb+1) -> (y, log|dy/dx|)."""
    inside = (x >= -bound) & (x <= bound)
    xs = torch.where(inside, x, torch.zeros_like(x))
    kx, w = _cumulative(uw, bound)
    ky, h = _cumulative(uh, bound)
    d = MIN_DERIV + F.softplus(ud)
    knots = ky if inverse else kx
    last = torch.zeros_like(knots)
    last[..., -1] = 1e-6
    nb = uw.shape[-1]
    k = ((xs[..., None] >= knots + last).sum(-1) - 1).clamp(0, nb - 1)[..., None]
    pick = lambda t: t.gather(-1, k)[..., 0]
    x0, w0, y0, h0 = pick(kx), pick(w), pick(ky), pick(h)
    s = pick(h / w)
    d0, d1 = pick(d), pick(d[..., 1:])
    t = d0 + d1 - 2 * s
    if inverse:
        dy = xs - y0
        a = dy * t + h0 * (s - d0)
        b = h0 * d0 - dy * t
        c = -s * dy
        root = (2 * c) / (-b - torch.sqrt((b * b - 4 * a * c).abs()))
        y = root * w0 + x0
        tt = root * (1 - root)
        den = s + t * tt
        num = s * s * (d1 * root * root + 2 * s * tt + d0 * (1 - root) ** 2)
        ld = -(torch.log(num) - 2 * torch.log(den))
    else:
        th = (xs - x0) / w0
        tt = th * (1 - th)
        den = s + t * tt
        y = y0 + h0 * (s * th * th + d0 * tt) / den
        num = s * s * (d1 * th * th + 2 * s * tt + d0 * (1 - th) ** 2)
        ld = torch.log(num) - 2 * torch.log(den)
    return torch.where(inside, y, x), torch.where(inside, ld, torch.zeros_like(ld))


class _FusedSplineFn(torch.autograd.Function):
    """Density-direction spline through fs_spline_train_fwd / fs_spline_train_bwd (one kernel each instead of ~80
    element-wise autograd kernels per direction).  theta: [B, N, 3 nb + 1] (conditional) or [N, 3 nb + 1] (parameters
    shared by all rows: the unconditional spline; its gradient is summed over the rows)."""

    @staticmethod
    def forward(ctx, x, theta, bound, nb, scale):
        from ._bridge import _lib
        x = x.contiguous()
        theta = theta.contiguous()
        B, N = x.shape
        shared = theta.dim() == 2
        y = torch.empty_like(x)
        ld = torch.empty_like(x)
        _lib.check(_lib.lib().fs_spline_train_fwd(_lib.ptr(x), _lib.ptr(theta), 0 if shared else N * (3 * nb + 1), B, N, nb,
                                                  float(bound), float(scale), _lib.ptr(y), _lib.ptr(ld),
                                                  _lib.stream_ptr(x.device)))
        ctx.save_for_backward(x, theta)
        ctx.meta = (float(bound), int(nb), float(scale), shared)
        return y, ld

    @staticmethod
    def backward(ctx, gy, gld):
        from ._bridge import _lib
        x, theta = ctx.saved_tensors
        bound, nb, scale, shared = ctx.meta
        B, N = x.shape
        gy = gy.contiguous()
        gld = gld.contiguous()
        gx = torch.empty_like(x)
        gt = torch.empty(B, N, 3 * nb + 1, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().fs_spline_train_bwd(_lib.ptr(x), _lib.ptr(theta), 0 if shared else N * (3 * nb + 1), B, N, nb,
                                                  bound, scale, _lib.ptr(gy), _lib.ptr(gld), _lib.ptr(gx), _lib.ptr(gt),
                                                  _lib.stream_ptr(x.device)))
        return gx, (gt.sum(0) if shared else gt), None, None, None


def fused_spline(x, theta, bound, nb, scale):
    """(y, log|dy/dx|) of the density-direction spline on CUDA float32 tensors; differentiable."""
    return _FusedSplineFn.apply(x, theta, bound, nb, scale)
