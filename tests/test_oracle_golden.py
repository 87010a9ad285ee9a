"""Pins the oracle (oracle/*.py) on outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py running
the reference's own code (the reference holds no golden vectors
for this path, so parity is pinned by executing it)."""
import os

import numpy as np
import pytest
import torch

from oracle import energy_ref as er
from oracle import flow_ref as fr
from oracle import mc_ref as mr

POT = er.Potential(2, [-10.0, -10.5], 1.2, 15.0)
NOPOT = er.Potential(0, [0, 0], 1.2, 15.0)


def _close(a, b, rtol):
    if np.isinf(b):
        return np.isinf(a) and a > 0
    return abs(a - b) <= rtol * max(1.0, abs(b))


def test_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "energy_cases.npz"))
    e, w = er.lj_energy_virial(g["kat_lj_r"])
    np.testing.assert_allclose(e, g["kat_lj_e"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(w, g["kat_lj_w"], rtol=1e-14, atol=1e-14)
    # Appendix B values, typed in independently of the fixture
    np.testing.assert_allclose(e[:3], [16128.016316891137, 0.016316891136000006, -0.983683108864], rtol=1e-13)
    assert e[4] == 0.0 and w[4] != 0.0 and e[5] == 0.0 and w[5] == 0.0
    v = er.double_well(g["kat_dw_pos"], 10, 10, POT)
    np.testing.assert_allclose(v, g["kat_dw_v"], rtol=1e-13, atol=1e-14)
    for pin, pout in zip(g["kat_pbc_in"], g["kat_pbc_out"]):
        np.testing.assert_array_equal(er.apply_pbc(pin, 10.0, 10.0), pout)
    assert er.apply_pbc(np.array([10.3, -1e-17]), 10.0, 10.0)[1] == 10.0


def test_energy_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "energy_cases.npz"))
    for name in g["names"]:
        p32 = g[name + "__pos"]
        L = float(g[name + "__L"])
        pot = POT if int(g[name + "__wells"]) else NOPOT
        for mode, rtol in (("f64", 1e-12), ("f32", 2e-6)):
            arr = p32.astype(np.float64) if mode == "f64" else p32
            E, W = er.total_energy_virial(arr, L, L, pot)
            assert _close(E, float(g["%s__%s_E" % (name, mode)]), rtol), (name, mode, E)
            assert _close(W, float(g["%s__%s_W" % (name, mode)]), rtol), (name, mode, W)
            for i, ref in zip(g[name + "__pidx"], g["%s__%s_pe" % (name, mode)]):
                e, w = er.particle_energy_virial(arr, int(i), L, L, pot)
                assert _close(e, ref[0], rtol) and _close(w, ref[1], rtol), (name, mode, i)
    assert float(g["kat3__f64_E"]) == pytest.approx(-21.123261258907107, rel=1e-6)   # float32 input rounding
    assert np.isinf(g["kat3_overlap__f64_E"]) and np.isinf(g["overlap_n32__f32_E"])


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_local_traces(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "mc_local.npz"))
    for key in g["names"]:
        if ("_%s_" % mode) not in key:
            continue
        p32 = g[key + "__pos0"]
        arr = p32.astype(np.float64) if mode == "f64" else p32.copy()
        L = float(g[key + "__L"])
        steps = int(g[key + "__steps"])
        # same PCG64 stream as the reference (monte_carlo.py:92-95)
        ch = mr.ChainRef(arr, L, 1.0, POT, max_displacement=float(g[key + "__md0"]),
                         rng=np.random.default_rng(int(g[key + "__seed"])))
        rtol = 1e-12 if mode == "f64" else 5e-6
        assert _close(ch.E, float(g[key + "__E0"]), rtol)
        for s in range(steps):
            p, eno, enn, acc, u = ch.local_step()
            assert p == int(g[key + "__idx"][s])
            assert _close(eno, g[key + "__eno"][s], rtol) and _close(enn, g[key + "__enn"][s], rtol), (key, s)
            assert int(acc) == int(g[key + "__acc"][s]), (key, s)
            if s + 1 == steps // 2:
                ch.adjust_displacement()
                assert ch.max_displacement == pytest.approx(float(g[key + "__md_mid"]), rel=1e-14)
        ch.adjust_displacement()
        assert ch.attempts == int(g[key + "__attempts"]) and ch.accepted == int(g[key + "__accepted"])
        assert ch.max_displacement == pytest.approx(float(g[key + "__mdF"]), rel=1e-14)
        np.testing.assert_allclose(np.asarray(ch.particles, np.float64), g[key + "__posF"],
                                   rtol=0, atol=1e-12 if mode == "f64" else 0)
        assert _close(ch.E, float(g[key + "__EF"]), 1e-11 if mode == "f64" else 5e-6)
        assert _close(ch.W, float(g[key + "__WF"]), 1e-11 if mode == "f64" else 5e-6)


def _load_sd(g, prefix="sd__"):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


@pytest.mark.parametrize("tag", ["n3_k3", "n4_k4", "n32_k2", "n4_k23", "n8_h128", "n6_h256"])
def test_flow(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "flow_%s.npz" % tag))
    sd = _load_sd(g)
    spec = fr.FlowSpec(sd, float(g["bound"]))
    assert spec.K == int(g["K"]) and spec.H == int(g["H"]) and spec.nb == int(g["nb"])
    assert spec.n_blocks == int(g["blocks"]) and spec.D == 2 * int(g["n"])
    x = torch.from_numpy(g["x"])
    z0 = torch.from_numpy(g["z0"])
    with torch.no_grad():
        y, ld = fr.layer_inverse(sd, spec.K - 1, spec, x)
        np.testing.assert_allclose(y.numpy(), g["lastlayer_inv"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ld.numpy(), g["lastlayer_ld"], rtol=1e-4, atol=1e-5)
        z, ldi = fr.inverse_and_log_det(sd, spec, x)
        np.testing.assert_allclose(z.numpy(), g["inv_z"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(ldi.numpy(), g["inv_ld"], rtol=1e-4, atol=1e-4)
        lp = fr.log_prob(sd, spec, x).numpy()
        ref = g["log_prob"]
        assert np.array_equal(np.isinf(lp), np.isinf(ref)) and np.isinf(ref).sum() == 1
        fin = ~np.isinf(ref)
        np.testing.assert_allclose(lp[fin], ref[fin], rtol=1e-5)
        xf, ldf = fr.forward_and_log_det(sd, spec, z0)
        np.testing.assert_allclose(xf.numpy(), g["fwd_x"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(ldf.numpy(), g["fwd_ld"], rtol=1e-4, atol=1e-4)
        # the reference's own round-trip property (flows/flow_test.py:40-48)
        xr, ldr = fr.forward_and_log_det(sd, spec, z)
        ok = (x.abs() <= spec.bound).all(dim=1)
        np.testing.assert_allclose(xr[ok].numpy(), x[ok].numpy(), atol=2e-3 * spec.bound)
        np.testing.assert_allclose((ldr + ldi)[ok].numpy(), 0, atol=5e-3)


def test_global_traces(golden_dir):
    g = np.load(os.path.join(golden_dir, "mc_global.npz"))
    for key in g["names"]:
        tag = key.split("_")[0]
        sd = _load_sd(g, tag + "__sd__")
        bound = float(g[tag + "__bound"])
        L = float(g[tag + "__L"])
        spec = fr.FlowSpec(sd, bound)
        ch = mr.ChainRef(g[key + "__pos0"].astype(np.float64), L, 1.0, POT, 0.65,
                         rng=np.random.default_rng(int(g[key + "__seed"])))
        for r in range(int(g[key + "__rounds"])):
            for _ in range(int(g[key + "__local"])):
                ch.local_step()
            cfg = g[key + "__props"][r]
            np.testing.assert_allclose(np.asarray(ch.particles, np.float64), g[key + "__pos_before"][r], atol=1e-12)
            assert ch.E == pytest.approx(float(g[key + "__eno"][r]), rel=1e-10, abs=1e-10)
            with torch.no_grad():
                old = torch.tensor((np.asarray(ch.particles) - L / 2).reshape(1, -1), dtype=torch.float)
                new = torch.tensor((cfg - np.array([L / 2, L / 2])).reshape(1, -1), dtype=torch.float)
                nll_old = -fr.log_prob(sd, spec, old).item()
                nll_new = -fr.log_prob(sd, spec, new).item()
            acc, _, _ = ch.global_move(cfg, nll_old, nll_new)
            assert int(acc) == int(g[key + "__acc"][r]), (key, r)
            e_after = float(g[key + "__E_after"][r])
            assert _close(ch.E, e_after, 1e-6), (key, r, ch.E, e_after)
        assert ch.attempts == int(g[key + "__attempts"]) and ch.accepted == int(g[key + "__accepted"])
        np.testing.assert_allclose(np.asarray(ch.particles, np.float64), g[key + "__posF"], atol=1e-6)


def test_observables_oracle_matches_reference_golden(golden_dir):
    """oracle/observables_ref.py against the outputs of the unmodified hybrid_NF_MCMC/utils.py functions."""
    from oracle import observables_ref as obr
    g = np.load(os.path.join(golden_dir, "observables.npz"))
    half_box, r0, start = float(g["ws_half_box"]), float(g["ws_r0"]), int(g["ws_start"])
    assert np.array_equal(obr.classify(g["ws_cfgs"], half_box, r0), g["ws_class"])
    avg_x, p_a, p_b, dF, runs = obr.well_statistics(g["ws_cfgs"], start, half_box, r0)
    assert np.array_equal(np.array(avg_x, dtype=np.float64), g["ws_avg_x"])
    assert np.array_equal(p_a, g["ws_p_a"]) and np.array_equal(p_b, g["ws_p_b"])
    np.testing.assert_allclose(dF, g["ws_dF"], rtol=0, atol=1e-15)
    for tag in "abc":
        r, gr = obr.pair_correlation(g["pc_%s_samples" % tag], int(g["pc_%s_n" % tag]), float(g["pc_%s_bound" % tag]),
                                     float(g["pc_%s_dr" % tag]))
        assert np.array_equal(r, g["pc_%s_r" % tag])
        np.testing.assert_allclose(gr, g["pc_%s_g" % tag], rtol=1e-14, atol=0)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (counter, key -> output), and the word -> draw mapping of the device."""
    from oracle import philox_ref as pr
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        got = pr.philox4x32(*[[c] for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == out
    p, u = pr.step_draws(0xa4093822 | (0x299f31d0 << 32), 0x03707344_13198a2e, 0x85a308d3_243f6a88, 1, 32)
    assert int(p[0]) == (0xd16cfe09 * 32) >> 32
    assert u[0].tolist() == [0x94fdcceb / 2 ** 32, 0x5001e420 / 2 ** 32, 0x24126ea1 / 2 ** 32]
    rng = pr.StepRNG(7, 3, 10, 4, 5)
    for s in range(4):
        assert rng.integers(5) == int(pr.step_draws(7, 3, 10 + s, 1, 5)[0][0])
        assert rng.random(2).shape == (2,)
    assert 0.0 <= rng.random() < 1.0


def test_target_energy_oracle_matches_reference_golden(golden_dir):
    """oracle/target_ref.py against NF.Energy.DoubleWellLJ._energy of the reference and its autograd gradient."""
    from oracle import target_ref as tr
    g = np.load(os.path.join(golden_dir, "target_energy.npz"))
    for tag in g["names"]:
        x = torch.from_numpy(g[tag + "__x"])
        n, b, T = int(g[tag + "__n"]), float(g[tag + "__bound"]), float(g[tag + "__T"])
        E32 = tr.double_well_lj_energy(x, n, T, b, [-10.0, -10.5], 1.2, 15).numpy()
        np.testing.assert_allclose(E32, g[tag + "__E"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(tr.simple_lj_energy(x, n, T, b).numpy(), g[tag + "__lj"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(tr.double_well(x, n, b, [-10.0, -10.5], 1.2, 15).numpy(), g[tag + "__dw"],
                                   rtol=2e-6, atol=2e-6)
        _, grad = tr.energy_and_grad(x, n, T, b, [-10.0, -10.5], 1.2, 15, dtype=torch.float32)
        np.testing.assert_allclose(grad.numpy(), g[tag + "__grad"], rtol=1e-4, atol=1e-4)


def test_affine_oracle_matches_reference_golden(golden_dir):
    """oracle/affine_ref.py against layer-by-layer outputs of the reference's MaskedAffineFlow / PeriodicShift stack,
    AffineCouplingBlocks (three scale maps, two split modes) and PeriodicWrap."""
    from oracle import affine_ref as ar
    g = np.load(os.path.join(golden_dir, "affine.npz"))
    D, bound = int(g["D"]), float(g["bound"])
    sd = {k[len("stack_sd__"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("stack_sd__")}
    z = torch.from_numpy(g["stack_z"])
    for i in range(8):
        if i % 2 == 0:
            b = sd["flows.%d.b" % i][0]
            s = ar.mlp(sd, "flows.%d.s." % i, b * z)
            t = ar.mlp(sd, "flows.%d.t." % i, b * z)
            z, ld = ar.masked_affine(z, b, s, t, False)
            np.testing.assert_allclose(ld.numpy(), g["stack_fwd_ld"][i], rtol=1e-5, atol=1e-6)
        else:
            z = ar.periodic_shift(z, list(range(0, D, 3)), bound, 0.37 * (i // 2 + 1))
        np.testing.assert_allclose(z.numpy(), g["stack_fwd_%d" % i], rtol=1e-5, atol=1e-5)
    x = torch.from_numpy(g["stack_x"])
    tot = torch.zeros(len(x))
    for i in range(7, -1, -1):
        if i % 2 == 0:
            b = sd["flows.%d.b" % i][0]
            x, ld = ar.masked_affine(x, b, ar.mlp(sd, "flows.%d.s." % i, b * x), ar.mlp(sd, "flows.%d.t." % i, b * x), True)
            tot = tot + ld
        else:
            x = ar.periodic_shift(x, list(range(0, D, 3)), bound, -0.37 * (i // 2 + 1))
        np.testing.assert_allclose(x.numpy(), g["stack_inv_%d" % i], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tot.numpy(), g["stack_inv_total_ld"], rtol=1e-5, atol=1e-5)
    for sm in ("exp", "sigmoid", "sigmoid_inv"):
        for mode in ("channel", "channel_inv"):
            tag = "blk_%s_%s" % (sm, mode)
            bsd = {k[len(tag + "_sd__"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "_sd__")}
            zz = torch.from_numpy(g["stack_z"])
            a, c = zz.chunk(2, dim=1)
            z1, z2 = (a, c) if mode == "channel" else (c, a)
            out, ld = ar.affine_coupling(z2, ar.mlp(bsd, "flows.1.param_map.", z1), sm, False)
            y = torch.cat([z1, out] if mode == "channel" else [out, z1], 1)
            np.testing.assert_allclose(y.numpy(), g[tag + "_fwd"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(ld.numpy(), g[tag + "_fwd_ld"], rtol=1e-5, atol=1e-5)
    far = torch.from_numpy(g["wrap_in"])
    np.testing.assert_allclose(ar.periodic_shift(far, list(range(1, D, 2)), bound, 0.0).numpy(), g["wrap_inv"],
                               rtol=1e-6, atol=1e-6)
