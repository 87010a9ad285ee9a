"""GPU parity: pair-energy kernels (through the C ABI) against the oracle and the
reference-generated golden vectors.  Tolerance: 1e-5 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import energy_ref as er

POT = er.Potential(2, [-10.0, -10.5], 1.2, 15.0)
NOPOT = er.Potential(0, [0, 0], 1.2, 15.0)
RTOL = 1e-5


def _mc():
    import flowstate_b200.MCMC as MC
    return MC


def _engine(pos, L, pot=POT):
    MC = _mc()
    return MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, pos.shape[1], num_wells=pot.num_wells,
                                V0_list=pot.V0_list, r0=pot.r0, k=pot.k, rng="philox")


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def test_known_answers(golden_dir):
    MC = _mc()
    g = np.load(os.path.join(golden_dir, "energy_cases.npz"))
    e, w = MC.lennard_jones_energy_virial(g["kat_lj_r"])
    # 2.5000001 is 2.5 in float32 (the device dtype): that entry sits ON the cut-off there
    keep = np.array([0, 1, 2, 3, 4, 6])
    np.testing.assert_allclose(e[keep], g["kat_lj_e"][keep], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(w[keep], g["kat_lj_w"][keep], rtol=2e-6, atol=2e-6)
    eo, wo = er.lj_energy_virial(g["kat_lj_r"].astype(np.float32))
    np.testing.assert_allclose(e, eo, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(w, wo, rtol=2e-6, atol=2e-6)
    assert abs(e[4]) < 1e-7 and w[4] != 0.0 and e[6] == 0.0 and w[6] == 0.0
    v = MC.double_well_potential(g["kat_dw_pos"], 10, 10, [-10, -10.5], 1.2, 15, 2)
    np.testing.assert_allclose(v, g["kat_dw_v"], rtol=1e-5, atol=1e-5)
    box = MC.SimulationBox(10.0)
    out = box.apply_pbc(np.array([-0.1, 10.0]))
    np.testing.assert_allclose(out, [9.9, 0.0], atol=1e-6)
    d = box.compute_distances(np.array([0.2, 0.3]), np.array([[9.7, 9.9], [5.0, 5.0]]))
    np.testing.assert_allclose(d[0], 0.6403124237432851, rtol=1e-6)


def test_golden_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "energy_cases.npz"))
    MC = _mc()
    for name in g["names"]:
        p32 = g[name + "__pos"]
        L = float(g[name + "__L"])
        pot = POT if int(g[name + "__wells"]) else NOPOT
        eng = _engine(p32[None], L, pot)
        E, W, ov = eng.total_energy_virial()
        E, W = float(E.item()), float(W.item())
        for mode in ("f64", "f32"):
            Er, Wr = float(g["%s__%s_E" % (name, mode)]), float(g["%s__%s_W" % (name, mode)])
            if np.isinf(Er):
                assert np.isinf(E) and E > 0 and np.isinf(W) and int(ov.item()) == 1, name
            else:
                assert _rel(E, Er) < RTOL and _rel(W, Wr) < RTOL, (name, mode, E, Er, W, Wr)
        ec = MC.EnergyCalculator(len(p32), p32, MC.SimulationBox(L), num_wells=pot.num_wells, V0_list=pot.V0_list,
                                 r0=pot.r0, k=pot.k, timing=False)
        assert ec.total_energy == pytest.approx(E, rel=1e-6) or (np.isinf(E) and np.isinf(ec.total_energy))
        for i, ref in zip(g[name + "__pidx"], g["%s__f64_pe" % name]):
            e, w = ec.calculate_particle_energy_virial(p32, int(i))
            if np.isinf(ref[0]):
                assert np.isinf(e) and np.isinf(w)
            else:
                assert _rel(e, ref[0]) < RTOL and _rel(w, ref[1]) < RTOL, (name, i)


@pytest.mark.parametrize("n,rho", [(3, 0.03), (5, 0.2), (32, 0.5), (33, 0.5), (64, 0.5), (100, 0.4), (256, 0.5),
                                   (257, 0.5), (1000, 0.5)])
def test_batches_against_oracle(n, rho):
    B = 6 if n >= 256 else 24
    pos, L = er.batch_lattices(B, n, rho, seed0=17 * n)
    pos[1, 0] = pos[1, n - 1] + np.float32(0.3)          # one overlapping configuration
    eng = _engine(pos, L)
    E, W, ov = eng.total_energy_virial()
    E, W, ov = E.cpu().numpy(), W.cpu().numpy(), ov.cpu().numpy()
    for b in range(B):
        Er, Wr = er.total_energy_virial(pos[b].astype(np.float64), L, L, POT)
        if np.isinf(Er):
            assert np.isinf(E[b]) and ov[b] == 1
        else:
            assert ov[b] == 0
            assert _rel(float(E[b]), Er) < RTOL and _rel(float(W[b]), Wr) < RTOL, (n, b, E[b], Er)
    assert ov[1] == 1


def test_particle_energy_against_oracle():
    import flowstate_b200._lib as lib
    n, B = 64, 16
    pos, L = er.batch_lattices(B, n, 0.5, seed0=5)
    dev = torch.device("cuda")
    tp = torch.from_numpy(pos).to(dev)
    idx = torch.arange(B, dtype=torch.int32, device=dev) * 3 % n
    new_xy = torch.from_numpy(pos[np.arange(B), (np.arange(B) * 3) % n] + np.float32(0.21)).to(dev).contiguous()
    pot = lib.make_pot(2, [-10, -10.5], 1.2, 15)
    for moved in (False, True):
        e = torch.empty(B, device=dev)
        w = torch.empty(B, device=dev)
        ov = torch.empty(B, dtype=torch.uint8, device=dev)
        lib.check(lib.lib().fs_energy_particle(lib.ptr(tp), lib.ptr(idx), lib.ptr(new_xy) if moved else None, B, n,
                                               L, L, pot, lib.ptr(e), lib.ptr(w), lib.ptr(ov), lib.stream_ptr()))
        for b in range(B):
            p = pos[b].astype(np.float64).copy()
            i = (b * 3) % n
            if moved:
                p[i] = new_xy[b].cpu().numpy().astype(np.float64)
            er_, wr_ = er.particle_energy_virial(p, i, L, L, POT)
            if np.isinf(er_):
                assert np.isinf(e[b].item())
            else:
                assert _rel(e[b].item(), er_) < RTOL and _rel(w[b].item(), wr_) < RTOL


def test_full_size_properties():
    """N=4096 (BASELINE config 5 upper end): properties that need no oracle."""
    n, B, rho = 4096, 8, 0.5
    pos, L = er.batch_lattices(B, n, rho, seed0=1)
    eng = _engine(pos, L)
    E, W, ov = eng.total_energy_virial()
    assert int(ov.sum().item()) == 0 and torch.isfinite(E).all()
    # relabelling particles leaves the total unchanged
    perm = np.random.default_rng(0).permutation(n)
    Ep, Wp, _ = eng.total_energy_virial(torch.from_numpy(np.ascontiguousarray(pos[:, perm])).cuda())
    torch.testing.assert_close(Ep, E, rtol=RTOL, atol=0)
    torch.testing.assert_close(Wp, W, rtol=RTOL, atol=0)
    # pair part is invariant under a rigid periodic shift (wells off)
    eng0 = _engine(pos, L, NOPOT)
    E0, _, _ = eng0.total_energy_virial()
    shifted = np.mod(pos + np.float32(L * 0.37), np.float32(L)).astype(np.float32)
    E1, _, _ = eng0.total_energy_virial(torch.from_numpy(shifted).cuda())
    torch.testing.assert_close(E1, E0, rtol=5e-5, atol=0)
    # one row against the oracle's single row sum (cheap slice of the O(N^2) work)
    r = er.distances(pos[0, 0].astype(np.float64), pos[0, 1:].astype(np.float64), L, L)
    e, _ = er.lj_energy_virial(r)
    import flowstate_b200._lib as lib
    idx = torch.zeros(B, dtype=torch.int32, device="cuda")
    pe = torch.empty(B, device="cuda")
    pw = torch.empty(B, device="cuda")
    lib.check(lib.lib().fs_energy_particle(lib.ptr(torch.from_numpy(pos).cuda()), lib.ptr(idx), None, B, n, L, L,
                                           lib.make_pot(0, [0, 0], 1.2, 15), lib.ptr(pe), lib.ptr(pw), None,
                                           lib.stream_ptr()))
    assert _rel(pe[0].item(), float(e.sum())) < RTOL


def test_empty_batch_and_bad_shapes():
    import flowstate_b200._lib as lib
    pot = lib.make_pot(2, [-10, -10.5], 1.2, 15)
    t = torch.zeros(1, device="cuda")
    assert lib.lib().fs_energy_total(lib.ptr(t), 0, 3, 10.0, 10.0, pot, lib.ptr(t), lib.ptr(t), None, None) == 0
    MC = _mc()
    with pytest.raises(ValueError):
        MC.BatchedMonteCarlo(np.zeros((2, 4, 2), np.float32), MC.SimulationBox(10.0), 1.0, 3)


@pytest.mark.parametrize("n", [32, 64, 256, 1000, 4096])
def test_fold_and_rint_minimum_image_agree(n):
    """Configurations inside [0, L] take the fold form of the minimum image (min(|d|, L - |d|)), a configuration with a
    particle outside the box the general d - L rint(d / L): the same configuration moved by one box length in one
    particle's coordinate must give the same energy (the shift is exact in float32 only up to the rounding of
    x - L, hence the tolerance, stated relative to the sum of the pair-term magnitudes)."""
    B = 32 if n < 1000 else (6 if n < 4096 else 2)
    pos, L = er.batch_lattices(B, n, 0.5, seed0=7 * n)
    eng = _engine(pos, L)
    E0, W0, _ = eng.total_energy_virial()
    sh = pos.copy()
    sh[:, 0, 0] -= np.float32(L)
    sh[:, n // 2, 1] += np.float32(L)
    E1, W1, _ = eng.total_energy_virial(torch.from_numpy(sh).cuda())
    for b in range(B):
        p64 = pos[b].astype(np.float64)
        scale = 0.5 * sum(er.particle_energy_magnitude(p64, p, L, L, POT) for p in range(n))
        assert abs(float(E0[b]) - float(E1[b])) <= 2e-5 * max(1.0, scale), (b, float(E0[b]), float(E1[b]), scale)
        Er, Wr = er.total_energy_virial(p64, L, L, POT)
        assert abs(float(E0[b]) - Er) <= RTOL * max(1.0, scale) and abs(float(E1[b]) - Er) <= 2e-5 * max(1.0, scale)
