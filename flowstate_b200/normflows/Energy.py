"""Base distribution of the flow: NF/normflows/Energy/Uniform.py:4-74."""
import torch
import torch.nn as nn


class UniformParticle(nn.Module):
    def __init__(self, n_particles, n_dimension, bound, device="cpu"):
        super().__init__()
        self.n_particles = n_particles
        self.n_dimension = n_dimension
        self.bound = bound
        self.device = device

    def sample(self, n_sample):
        z = torch.empty((n_sample, self.n_particles, self.n_dimension), dtype=torch.float32,
                        device=self.device).uniform_(-self.bound, self.bound)
        return z.reshape(n_sample, self.n_particles * self.n_dimension)

    def forward(self, n_sample):
        return self.sample(n_sample)

    def log_prob(self, z):
        inside = ((z >= -self.bound) & (z <= self.bound)).all(dim=1)
        d = self.n_particles * self.n_dimension
        c = -d * torch.log(torch.tensor(2 * self.bound))
        out = torch.full((z.size(0),), float(c), device=z.device, dtype=z.dtype)
        out[~inside] = -float("inf")
        return out


class _TargetEnergyFn(torch.autograd.Function):
    """fs_target_energy with its analytic gradient (one kernel produces both)."""

    @staticmethod
    def forward(ctx, x, mod):
        E, g = mod._launch(x, True)
        ctx.save_for_backward(g)
        return E

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return grad_out[:, None] * g, None


class SimpleLJ(nn.Module):
    """NF/normflows/Energy/SimpleLJ.py:5-39: soft-core Lennard-Jones target of reverse_kld (an extra particle at the
    origin, coordinates wrapped into the box, no minimum image, divided by the temperature).  Evaluated by the CUDA
    kernel fs_target_energy (value and gradient); there is no CPU fallback - the reference itself allocates on 'cuda'
    unconditionally (SimpleLJ.py:21)."""

    def __init__(self, dim, n_particles, temperature, bound):
        super().__init__()
        self._dim = dim
        self._n_particles = n_particles
        self._n_dimensions = dim // n_particles
        self.temperature = temperature
        self.bound = bound

    def _pot(self, wells=True):
        from ._bridge import _lib
        return _lib.make_pot(0, [0.0, 0.0], 0.0, 0.0)

    def _launch(self, x, want_grad, wells=True):
        from ._bridge import _lib
        if self._n_dimensions != 2:
            raise _lib.FlowStateError("flowstate_b200: the target energy kernel is two-dimensional")
        xc = _lib.require_cuda(x.detach().contiguous().float(), "target energy input")
        B = xc.shape[0]
        E = torch.empty(B, dtype=torch.float32, device=xc.device)
        g = torch.empty_like(xc) if want_grad else None
        if B == 0:
            return E, g
        pot = self._pot(wells)
        _lib.check(_lib.lib().fs_target_energy(_lib.ptr(xc), B, self._n_particles, float(self.bound),
                                               float(self.temperature), pot, _lib.ptr(E), _lib.ptr(g),
                                               _lib.stream_ptr(xc.device)))
        return E, g

    def _energy(self, x):
        x = x.reshape(x.shape[0], self._dim)
        if x.requires_grad and torch.is_grad_enabled():
            return _TargetEnergyFn.apply(x, self)
        return self._launch(x, False)[0]



class DoubleWellLJ(SimpleLJ):
    """NF/normflows/Energy/SimpleLJ.py:42-128: SimpleLJ plus the tanh double well with centres (-bound/2, 0) and
    (bound/2, 0) (minimum image, not divided by the temperature); main_algorithm_2.py:282-285."""

    def __init__(self, dim, n_particles, temperature, bound, V0_list=None, r0=1.0, k=10.0):
        super().__init__(dim, n_particles, temperature, bound)
        if V0_list is None:
            V0_list = [-4.0, -4.0]
        self.V0_list = torch.tensor(V0_list, dtype=torch.float32)
        self.r0 = r0
        self.k = k
        self.centers = torch.tensor([[-bound / 2, 0.0], [bound / 2, 0.0]], dtype=torch.float32)

    def _pot(self, wells=True):
        from ._bridge import _lib
        if not wells:
            return SimpleLJ._pot(self)
        return _lib.make_pot(2, [float(v) for v in self.V0_list], self.r0, self.k)

    def double_well_potential(self, positions):
        """SimpleLJ.py:61-112 alone (wells only), summed over the particles of each configuration."""
        x = positions.reshape(positions.shape[0], self._dim)
        return self._launch(x, False)[0] - self._launch(x, False, wells=False)[0]
