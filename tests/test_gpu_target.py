"""GPU parity of the Algorithm-2 training target (fs_target_energy <- NF.Energy.DoubleWellLJ._energy,
NF/normflows/Energy/SimpleLJ.py:15-128) and of NormalizingFlow.reverse_kld (NF/normflows/core.py:110-141), which the
reference's unmodified Alg-2 driver calls every batch (main_algorithm_2.py:319, 446)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import target_ref as tr

V0, R0, K = [-10.0, -10.5], 1.2, 15


def _target(n, T, bound):
    import flowstate_b200.normflows as NF
    return NF.Energy.DoubleWellLJ(2 * n, n, T, bound, V0_list=V0, r0=R0, k=K)


def test_golden_energy_and_gradient(golden_dir):
    g = np.load(os.path.join(golden_dir, "target_energy.npz"))
    for tag in g["names"]:
        n, b, T = int(g[tag + "__n"]), float(g[tag + "__bound"]), float(g[tag + "__T"])
        tgt = _target(n, T, b)
        x = torch.from_numpy(g[tag + "__x"]).cuda().requires_grad_(True)
        E = tgt._energy(x)
        (grad,) = torch.autograd.grad(E.sum(), x)
        # float64 oracle = truth; the reference's own float32 result (golden) is 4e-6 away from it
        E64, g64 = tr.energy_and_grad(x.detach().cpu(), n, T, b, V0, R0, K)
        errE = np.max(np.abs(E.detach().cpu().numpy() - E64.numpy()) / np.maximum(1.0, np.abs(E64.numpy())))
        errG = np.max(np.abs(grad.cpu().numpy() - g64.numpy()) / np.maximum(1.0, np.abs(g64.numpy())))
        print("%s: energy err %.2e, gradient err %.2e vs float64 oracle" % (tag, errE, errG))
        assert errE < 1e-5 and errG < 2e-5
        np.testing.assert_allclose(E.detach().cpu().numpy(), g[tag + "__E"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(grad.cpu().numpy(), g[tag + "__grad"], rtol=1e-4, atol=1e-4)
        # parts: SimpleLJ alone, wells alone; no-grad path gives the same values
        import flowstate_b200.normflows as NF
        lj = NF.Energy.SimpleLJ(2 * n, n, T, b)._energy(x.detach())
        np.testing.assert_allclose(lj.cpu().numpy(), g[tag + "__lj"], rtol=1e-5, atol=1e-5)
        dw = tgt.double_well_potential(x.detach().view(-1, n, 2))
        np.testing.assert_allclose(dw.cpu().numpy(), g[tag + "__dw"], rtol=1e-5, atol=2e-5)
        with torch.no_grad():
            assert torch.equal(tgt._energy(x), E.detach())


def test_gradient_scaling_and_empty_batch():
    n, b = 8, 6.0
    tgt = _target(n, 1.3, b)
    x = ((torch.rand(16, 2 * n) * 2 - 1) * b).cuda().requires_grad_(True)
    w = torch.linspace(0.5, 2.0, 16).cuda()
    (g1,) = torch.autograd.grad((tgt._energy(x) * w).sum(), x)
    _, g64 = tr.energy_and_grad(x.detach().cpu(), n, 1.3, b, V0, R0, K)
    ref = g64.numpy() * w.cpu().numpy()[:, None]
    assert np.max(np.abs(g1.cpu().numpy() - ref) / np.maximum(1.0, np.abs(ref))) < 2e-5
    assert tgt._energy(torch.empty(0, 2 * n, device="cuda")).shape == (0,)


def test_reverse_kld_runs_like_the_reference_driver():
    """main_algorithm_2.py:446-451: energy_loss, z = model.reverse_kld(BATCH_SIZE); loss.backward() reaches every
    flow parameter; the energy term equals the oracle's on the returned z."""
    import flowstate_b200.normflows as NF
    torch.manual_seed(0)
    n, bound = 6, 5.0
    base = NF.Energy.UniformParticle(n, 2, bound, device="cuda")
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, 2, 64, range(2 * n), num_bins=8, tail_bound=bound)
              for _ in range(3)]
    model = NF.NormalizingFlow(base, layers, _target(n, 1.0, bound)).cuda()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    model.train()
    loss, z = model.reverse_kld(64)
    assert z.shape == (64, 2 * n) and torch.isfinite(loss)
    loss.backward()
    grads = [p.grad for name, p in model.named_parameters() if "preprocessing" not in name]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) for g in grads) > 0
    e = model.p._energy(z.detach())
    e_ref = tr.double_well_lj_energy(z.detach().cpu().double(), n, 1.0, bound, V0, R0, K)
    assert np.max(np.abs(e.cpu().numpy() - e_ref.numpy()) / np.maximum(1.0, np.abs(e_ref.numpy()))) < 1e-5
