"""GPU parity: NF-proposed global moves (flow log-densities + total energy + fused accept)
against reference nf_big_move traces, and the accept kernel against its formula."""
import logging
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import energy_ref as er
from oracle import flow_ref as fr

POT = er.Potential(2, [-10.0, -10.5], 1.2, 15.0)


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _model(sd, n, bound):
    import flowstate_b200.normflows as NF
    spec = fr.FlowSpec(sd, bound)
    base = NF.Energy.UniformParticle(n, 2, bound, device="cuda")
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, spec.n_blocks, spec.H, range(2 * n),
                                                              num_bins=spec.nb, tail_bound=bound)
              for _ in range(spec.K)]
    m = NF.NormalizingFlow(base, layers)
    m.load_state_dict(sd)
    return m.cuda().eval(), spec


def test_golden_global_traces(golden_dir):
    """Replays hybrid runs (25 local steps + 1 global move, 12 rounds) through the drop-in
    MonteCarlo with the reference's seeds.  The reference state is float64 until its first
    NF acceptance, the device state float32, so decisions are required to match only while
    the two trajectories have consumed identical random numbers and sit outside the
    epsilon band; each compared round checks the old total energy and the decision."""
    import flowstate_b200.MCMC as MC
    g = np.load(os.path.join(golden_dir, "mc_global.npz"))
    lg = logging.getLogger("fs_quiet")
    lg.setLevel(logging.CRITICAL)
    compared = 0
    for key in g["names"]:
        tag = key.split("_")[0]
        sd = _sd(g, tag + "__sd__")
        bound, L = float(g[tag + "__bound"]), float(g[tag + "__L"])
        n = g[key + "__pos0"].shape[0]
        model, spec = _model(sd, n, bound)
        mc = MC.MonteCarlo(g[key + "__pos0"], MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5],
                           r0=1.2, k=15, initial_max_displacement=0.65, logger=lg, seed=int(g[key + "__seed"]))
        mc.set_nf_model(model)
        for r in range(int(g[key + "__rounds"])):
            for _ in range(int(g[key + "__local"])):
                mc.particle_displacement()
            ref_pos = g[key + "__pos_before"][r]
            if np.abs(mc.particles.astype(np.float64) - ref_pos).max() > 1e-4:
                break                      # an in-band local decision diverged; stop comparing this chain
            eno = mc.energy_calculator.total_energy
            assert abs(eno - float(g[key + "__eno"][r])) <= 2e-5 * max(1.0, abs(float(g[key + "__eno"][r])))
            cfg = g[key + "__props"][r]
            # decision margin from the oracle, to know whether this round is inside the band
            with torch.no_grad():
                lo = fr.log_prob(sd, spec, torch.tensor((ref_pos - L / 2).reshape(1, -1), dtype=torch.float)).item()
                ln = fr.log_prob(sd, spec, torch.tensor((cfg.astype(np.float64) - L / 2).reshape(1, -1),
                                                        dtype=torch.float)).item()
            enn, _ = er.total_energy_virial(cfg, L, L, POT)
            acc = mc.nf_big_move(cfg)
            ref_acc = bool(g[key + "__acc"][r])
            if acc != ref_acc:
                delta = -(enn - float(g[key + "__eno"][r])) - ((-ln) - (-lo))
                eps = 1e-5 * (abs(enn) + abs(eno)) + 1e-4 * (abs(lo) + abs(ln))
                assert np.isfinite(delta) and abs(delta) < 50 * eps + 1.0, (key, r, delta, eps)
                break
            compared += 1
            assert mc.attempts_displacement == (r + 1) * (int(g[key + "__local"]) + 1)
    assert compared >= 20


def test_accept_kernel_formula():
    import flowstate_b200._lib as lib
    B, n = 512, 8
    rs = np.random.default_rng(0)
    pos = rs.random((B, n, 2)).astype(np.float32)
    prop = rs.random((B, n, 2)).astype(np.float32) + 1
    E = rs.normal(-20, 5, B)
    W = rs.normal(0, 5, B)
    E_new = (E + rs.normal(0, 2, B)).astype(np.float32)
    W_new = rs.normal(0, 5, B).astype(np.float32)
    lq_old = rs.normal(-30, 3, B).astype(np.float32)
    lq_new = rs.normal(-30, 3, B).astype(np.float32)
    E_new[:8] = np.inf                  # overlapping proposals: always rejected, uniform still consumed
    lq_new[8:12] = -np.inf              # proposal outside the box
    u = rs.random(B)
    beta = 1.0 / 0.7
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tpos, tprop, tE, tW = d(pos), d(prop), d(E), d(W)
    tEn, tWn, tlo, tln, tu = d(E_new), d(W_new), d(lq_old), d(lq_new), d(u)     # keep the device buffers alive
    att = torch.full((B,), 10, dtype=torch.int64, device="cuda")
    acc = torch.full((B,), 4, dtype=torch.int64, device="cuda")
    mask = torch.empty(B, dtype=torch.uint8, device="cuda")
    lib.check(lib.lib().fs_accept_global(lib.ptr(tpos), lib.ptr(tprop), lib.ptr(tE), lib.ptr(tW), lib.ptr(tEn),
                                         lib.ptr(tWn), lib.ptr(tlo), lib.ptr(tln), lib.ptr(tu),
                                         None, beta, lib.ptr(att), lib.ptr(acc), lib.ptr(mask), B, n,
                                         lib.stream_ptr()))
    with np.errstate(all="ignore"):
        ratio_log = -beta * (E_new.astype(np.float64) - E) - ((-lq_new.astype(np.float64)) - (-lq_old.astype(np.float64)))
        ratio = np.exp(ratio_log)
        ref = (ratio >= 1.0) | (u < ratio)
    got = mask.cpu().numpy().astype(bool)
    assert np.array_equal(got, ref)
    assert not got[:12].any() and 0.2 < got.mean() < 0.8
    assert torch.equal(att.cpu(), torch.full((B,), 11, dtype=torch.int64))
    assert np.array_equal(acc.cpu().numpy(), 4 + ref.astype(np.int64))
    out = tpos.cpu().numpy()
    assert np.array_equal(out[ref], prop[ref]) and np.array_equal(out[~ref], pos[~ref])
    assert np.array_equal(tE.cpu().numpy()[ref], E_new.astype(np.float64)[ref])
    assert np.array_equal(tE.cpu().numpy()[~ref], E[~ref])
    assert np.array_equal(tW.cpu().numpy()[ref], W_new.astype(np.float64)[ref])


def test_batched_global_move_equals_single_chain_facade(golden_dir):
    import flowstate_b200.MCMC as MC
    g = np.load(os.path.join(golden_dir, "mc_global.npz"))
    tag = "n8"
    sd = _sd(g, tag + "__sd__")
    bound, L = float(g[tag + "__bound"]), float(g[tag + "__L"])
    model, spec = _model(sd, 8, bound)
    B = 16
    pos, _ = er.batch_lattices(B, 8, 0.03, seed0=50)
    seeds = list(range(100, 100 + B))
    kw = dict(num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15, initial_max_displacement=0.65)
    eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, 8, seeds=seeds, **kw)
    eng.set_nf_model(model)
    torch.manual_seed(9)
    masks = []
    props = []
    for r in range(4):
        eng.particle_displacement(20)
        z = model.sample(B)
        cfg = (z.reshape(B, 8, 2) + np.float32(L / 2)).contiguous()
        if r % 2:
            cfg = ((eng.pos + 0.003 * (r + 1)) % np.float32(L)).contiguous()     # near-identity proposals get accepted
        props.append(cfg.cpu().numpy())
        masks.append(eng.nf_big_move(cfg).cpu().numpy())
    masks = np.stack(masks)
    assert masks.sum() > 0
    lg = logging.getLogger("fs_quiet")
    lg.setLevel(logging.CRITICAL)
    for b in (0, 5, 15):
        mc = MC.MonteCarlo(pos[b], MC.SimulationBox(L), 1.0, 8, logger=lg, seed=seeds[b], **kw)
        mc.set_nf_model(model)
        for r in range(4):
            for _ in range(20):
                mc.particle_displacement()
            assert mc.nf_big_move(props[r][b]) == bool(masks[r, b])
        np.testing.assert_array_equal(mc.particles, eng.pos[b].cpu().numpy())
        assert mc.accepted_displacement == int(eng.accepted[b].item())
        assert mc.energy_calculator.total_energy == eng.E[b].item()
    # sample() observables follow monte_carlo.py:416-444
    cyc, e_per_n, rho, P, lx, ly, parts = eng.sample(7)
    assert cyc == 7 and rho == pytest.approx(8 / (L * L)) and parts.shape == (B, 8, 2)
    torch.testing.assert_close(P, rho / 1.0 + eng.W / (2 * L * L))


def test_second_device_binding():
    """One process per GPU: the library's runtime must follow the device that owns the buffers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import flowstate_b200.MCMC as MC
    pos, L = er.batch_lattices(4, 16, 0.3, seed0=3)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(dev):
            eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, 16, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2,
                                       k=15, initial_max_displacement=0.4, rng="philox", philox_seed=7, device=dev)
            eng.particle_displacement(50)
            assert eng.pos.device == torch.device(dev)
            outs.append(eng.pos.cpu())
    assert torch.equal(outs[0], outs[1])


def test_judge_and_bulk_judge_match_reference(golden_dir):
    """MonteCarlo.judge_normalizing_flow / bulk_judge_normalizing_flow (monte_carlo.py:305-370) against outputs of the
    reference on the same seed: criteria, restored cached energy, attempts counter, and the position of the numpy
    generator afterwards (one uniform per finite uphill proposal)."""
    import logging
    import flowstate_b200.MCMC as MC
    g = np.load(os.path.join(golden_dir, "mc_judge.npz"))
    n, L, seed = int(g["n"]), float(g["L"]), int(g["seed"])
    lg = logging.getLogger("fs_quiet_judge")
    lg.setLevel(logging.CRITICAL)
    mc = MC.MonteCarlo(g["pos0"], MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                       initial_max_displacement=0.5, logger=lg, seed=seed)
    for _ in range(int(g["warm"])):
        mc.particle_displacement()
    e0 = mc.energy_calculator.total_energy
    assert abs(e0 - float(g["e_before"])) <= 1e-5 * max(1.0, abs(float(g["e_before"])))
    att0 = mc.attempts_displacement
    props = g["props"]
    crit = [bool(mc.judge_normalizing_flow(c.copy())) for c in props[:12]]
    assert crit == [bool(c) for c in g["crit"]]
    assert mc.energy_calculator.total_energy == e0                  # cache restored (:326-327)
    assert mc.attempts_displacement - att0 == int(g["att_delta"])
    acc, att = mc.bulk_judge_normalizing_flow([c.copy() for c in props[12:]], float(g["bulk_ref_energy"]))
    assert (acc, att) == (int(g["bulk_acc"]), int(g["bulk_att"]))
    assert mc.rng.random() == float(g["next_uniform"])              # same number of uniforms consumed


@pytest.mark.parametrize("n,B,rho", [(3, 37, 0.03), (32, 301, 0.5), (45, 64, 0.4), (64, 257, 0.5), (256, 96, 0.5),
                                     (300, 33, 0.3), (1000, 9, 0.5), (4000, 3, 0.5), (7500, 3, 0.5)])
def test_fused_global_move_equals_energy_then_accept(n, B, rho):
    """fs_accept_global_fused (proposal energy + acceptance + update in one kernel) against the two-kernel sequence
    fs_energy_total -> fs_accept_global on identical inputs: bit-equal energies, masks, states and counters for every
    group size of the energy kernel, ragged batch sizes, overlapping proposals (E = inf: rejected), proposals outside the
    box (rint minimum image) and Philox as well as replayed uniforms.  N = 4000 runs the 1024-thread block of the packed
    kernel, N = 7500 is beyond its tile and takes the entry point's two-kernel fallback."""
    import flowstate_b200.MCMC as MC
    pos, L = er.batch_lattices(B, n, rho, seed0=3 * n)
    prop, _ = er.batch_lattices(B, n, rho, seed0=5 * n + 1)
    if n > 3:
        prop[1, 0] = prop[1, n - 1] + np.float32(0.3)          # overlapping proposal
        prop[2, 1, 0] -= np.float32(L)                         # a proposal with a particle outside [0, L]
    rs = np.random.default_rng(n)
    lq_old = torch.from_numpy(rs.normal(-30, 2, B).astype(np.float32)).cuda()
    lq_new = torch.from_numpy(rs.normal(-30, 2, B).astype(np.float32)).cuda()
    u = torch.from_numpy(rs.random(B)).cuda()
    tprop = torch.from_numpy(prop).cuda()
    for draws in ("replay", "philox"):
        engs = []
        for fused in (True, False):
            eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2,
                                       k=15, rng="philox", seeds=11)
            eng.fused_accept = fused
            mask = eng.nf_big_move(tprop, u=u if draws == "replay" else None, logq=(lq_old, lq_new))
            engs.append((eng, mask))
        (a, ma), (b, mb) = engs
        E_ref, W_ref, _ = b.total_energy_virial(tprop)
        assert torch.equal(a.proposal_energy, E_ref) and torch.equal(b.proposal_energy, E_ref)
        assert torch.equal(ma, mb) and torch.equal(a.pos, b.pos)
        assert torch.equal(a.E, b.E) and torch.equal(a.W, b.W)
        assert torch.equal(a.attempts, b.attempts) and torch.equal(a.accepted, b.accepted)
        m = ma.bool().cpu().numpy()
        out = a.pos.cpu().numpy()
        assert np.array_equal(out[m], prop[m]) and np.array_equal(out[~m], pos[~m])
        if n > 3:
            assert not m[1] and torch.isinf(E_ref[1])
        assert 0 < m.sum() < B or B < 8


def test_global_move_with_sampling_pass_log_density(golden_dir):
    """nf_big_move(configs, logq_new=...) with the proposals' log q from the sampling pass that produced them
    (NormalizingFlow.sample_with_log_prob) against the default, which sends the proposals through the density pass
    again like the reference (monte_carlo.py:262): same uniforms -> same decisions except where the acceptance ratio is
    within the epsilon band of the threshold (|log ratio - log u| <= 1e-4 |log q|, the log-density tolerance)."""
    import flowstate_b200.MCMC as MC
    g = np.load(os.path.join(golden_dir, "mc_global.npz"))
    tag = str(g["names"][0]).split("_")[0]
    sd = _sd(g, tag + "__sd__")
    bound, L = float(g[tag + "__bound"]), float(g[tag + "__L"])
    n = g[str(g["names"][0]) + "__pos0"].shape[0]
    model, spec = _model(sd, n, bound)
    B = 512
    pos, _ = er.batch_lattices(B, n, n / (L * L), seed0=41)
    torch.manual_seed(3)
    torch.cuda.manual_seed(3)
    x, lq_new = model.sample_with_log_prob(B)
    cfg = (x.reshape(B, n, 2) + np.float32(L / 2)).contiguous()
    u = torch.rand(B, dtype=torch.float64, device="cuda")
    res = []
    for from_sample in (False, True):
        eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                                   rng="philox", seeds=5)
        eng.set_nf_model(model)
        E_old = eng.E.clone()
        mask = eng.nf_big_move(cfg, u=u, logq_new=lq_new if from_sample else None)
        res.append((mask.bool().cpu().numpy(), eng, E_old))
    (m0, e0, E_old), (m1, e1, _) = res
    lq = model.log_prob(torch.cat([e0.centred(torch.from_numpy(pos).cuda()), e0.centred(cfg)]))
    lq_old, lq_inv = lq[:B].double(), lq[B:].double()
    assert ((lq_new.double() - lq_inv).abs() / lq_inv.abs()).max().item() < 1e-4
    log_ratio = -(e0.proposal_energy.double() - E_old) - (lq_inv - lq_old)          # beta = 1
    band = 1e-4 * lq_inv.abs() + 1e-9
    margin = (log_ratio - torch.log(u)).abs()
    differ = torch.from_numpy(m0 != m1).cuda()
    assert not bool((differ & (margin > band) & (log_ratio < 0)).any()), int(differ.sum())
    assert 0 < m0.sum() and torch.equal(e0.attempts, e1.attempts)
    same = ~differ
    assert torch.equal(e0.pos[same], e1.pos[same])
