"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the Algorithm-2 training target.

Follows NF/normflows/Energy/SimpleLJ.py of the reference:
  :15-39   SimpleLJ._energy            wrap x - 2b round(x / 2b); prepend a particle at the origin; all pair distances
                                       WITHOUT minimum image; r <= 0.82 -> -80 (r - 0.82) + 30, else 4 (r^-12 - r^-6);
                                       sum over pairs, divided by the temperature
  :61-112  DoubleWellLJ.double_well_potential   centres (-b/2, 0), (b/2, 0), minimum image with L = 2b,
                                       V0_i (1 - 0.5 (1 + tanh(k (r - r0)))) summed over wells and particles
  :114-128 DoubleWellLJ._energy        LJ / T + wells
The reference allocates its zero particle on 'cuda' unconditionally (:21); this restatement runs on the CPU in the
requested dtype (float32 mirrors the reference's arithmetic, float64 is the truth used to attribute error) and is
differentiable by torch autograd, which gives the reference gradient of the reverse-KL energy term.
Pinned against reference outputs in tests/golden/target_energy.npz (oracle/make_golden.py: gen_target).
"""
import torch

BKPOINT = 0.82


def simple_lj_energy(x, n_particles, temperature, bound):
    x = x.reshape(x.shape[0], n_particles, -1)
    d = x - 2 * bound * torch.round(x / (bound * 2))
    d = torch.cat((torch.zeros(d.shape[0], 1, d.shape[2], dtype=d.dtype), d), dim=1)
    diff = d.unsqueeze(2) - d.unsqueeze(2).transpose(1, 2)
    iu = torch.triu_indices(n_particles + 1, n_particles + 1, offset=1)
    dv = diff[:, iu[0], iu[1], :]
    r = torch.sqrt((dv * dv).sum(-1))
    e = torch.where(r <= BKPOINT, -80 * (r - BKPOINT) + 30, 4 * (torch.pow(1 / r, 12) - torch.pow(1 / r, 6)))
    return e.sum(dim=1) / temperature


def double_well(x, n_particles, bound, V0_list, r0, k):
    pos = x.reshape(x.shape[0], n_particles, -1)
    L = 2 * bound
    V = torch.zeros(pos.shape[0], dtype=pos.dtype)
    for i, cx in enumerate((-bound / 2, bound / 2)):
        dx = pos[:, :, 0] - cx
        dy = pos[:, :, 1] - 0.0
        dx = dx - L * torch.round(dx / L)
        dy = dy - L * torch.round(dy / L)
        r = torch.sqrt(dx ** 2 + dy ** 2)
        V = V + (V0_list[i] * (1 - 0.5 * (1 + torch.tanh(k * (r - r0))))).sum(dim=1)
    return V


def double_well_lj_energy(x, n_particles, temperature, bound, V0_list, r0, k):
    return simple_lj_energy(x, n_particles, temperature, bound) + double_well(x, n_particles, bound, V0_list, r0, k)


def energy_and_grad(x, n_particles, temperature, bound, V0_list, r0, k, dtype=torch.float64):
    """(E [B], dE/dx [B, D]) by autograd, plus the magnitude sum |pair terms| + |well terms| the tolerance scales with."""
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    E = double_well_lj_energy(xx, n_particles, temperature, bound, V0_list, r0, k)
    (g,) = torch.autograd.grad(E.sum(), xx)
    return E.detach(), g
