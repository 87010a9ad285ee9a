"""GPU: the batched hybrid drivers run end to end on small settings (training = autograd path on the
device, sampling = CUDA kernels), and training actually lowers the forward-KL loss."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_algorithm_1_small():
    from flowstate_b200.drivers import HybridConfig, run_algorithm_1
    torch.manual_seed(0)
    cfg = HybridConfig(particles=3, chains=16, equilibration_steps=600, adjusting_frequency=300, sampling_frequency=20,
                       K=3, blocks=2, hidden=128, bins=8, lr=2e-3, batch_size=128, epochs=6, training_samples=1600,
                       big_move_attempts=10, big_move_interval=50)
    out = run_algorithm_1(cfg, log=lambda *a: None)
    eng = out["engine"]
    assert out["attempts"] == 16 * (600 + 100 * 20 + 10 * 51)
    assert 0 < out["accepted"] < out["attempts"]
    assert 0 <= out["big_move_accepts"] <= 160
    assert np.isfinite(out["final_loss"])
    # the flow learned something: its log-density of chain states beats the untrained (uniform) flow
    lq = out["model"].log_prob(eng.centred(eng.pos))
    assert torch.isfinite(lq).all()
    assert lq.mean().item() > -6 * np.log(10.0) + 0.5
    # running energies still agree with a recomputation
    E = eng.E.clone()
    eng.refresh_energy()
    assert ((E - eng.E).abs() / eng.E.abs().clamp(min=1)).max().item() < 1e-5


def test_algorithm_2_small():
    from flowstate_b200.drivers import HybridConfig, run_algorithm_2
    torch.manual_seed(0)
    cfg = HybridConfig.preset(2, particles=4, chains=32, equilibration_steps=300, K=2, blocks=2, hidden=128, bins=8,
                              lr=1e-3, batch_size=64, cycles=5, local_steps=40, training_samples=128)
    out = run_algorithm_2(cfg, log=lambda *a: None)
    # equilibration + initial training set (128 / (32 / 10) = 40 steps per chain) + 5 cycles of 40 local + 1 global move
    assert out["attempts"] == 32 * (300 + 40 + 5 * 41)
    assert np.isfinite(out["final_loss"])
    # with a reverse-KL share (ALPHA < 1, main_algorithm_2.py:446-448) the DoubleWellLJ target kernel is on the path
    cfg_r = HybridConfig.preset(2, particles=4, chains=32, equilibration_steps=100, K=2, blocks=2, hidden=128, bins=8,
                                lr=1e-3, batch_size=64, cycles=2, local_steps=40, training_samples=128, alpha=0.5)
    out_r = run_algorithm_2(cfg_r, log=lambda *a: None)
    assert np.isfinite(out_r["final_loss"])


def test_presets_match_the_reference_drivers():
    """--algorithm 2 with defaults builds the reference's Alg-2 flow (main_algorithm_2.py:33-76: K = 23, H = 128,
    2 blocks, 15 bins, Adam 5.435e-4 / 9.586e-5, batch 256) on the throughput RNG; Algorithm 1 keeps K = 15, H = 256,
    32 blocks (NUM_BINS in the num_blocks slot, main_algorithm_1.py:282), 32 bins."""
    from flowstate_b200.drivers import HybridConfig, hybrid
    c2 = HybridConfig.preset(2)
    assert (c2.K, c2.hidden, c2.blocks, c2.bins, c2.batch_size, c2.chains) == (23, 128, 2, 15, 256, 100)
    assert abs(c2.lr - 5.43510751759681e-4) < 1e-18 and abs(c2.weight_decay - 9.5857178422352e-05) < 1e-18
    assert c2.rng == "philox" and c2.precision == "auto" and c2.sampling_frequency == 10
    L = float(np.float32(np.sqrt(c2.particles / c2.rho)))
    m = hybrid._build_flow(c2, L, "cuda")
    net = m.flows[0].prqct.transform_net
    assert len(m.flows) == 23 and net.hidden_features == 128 and len(net.blocks) == 2 and m.flows[0].prqct.num_bins == 15
    assert type(m.p).__name__ == "DoubleWellLJ"
    c1 = HybridConfig.preset(1)
    assert (c1.K, c1.hidden, c1.blocks, c1.bins, c1.batch_size, c1.lr) == (15, 256, 32, 32, 512, 1e-4)
    c0 = HybridConfig.preset(0)
    assert c0.chains == 100 and c0.production_steps == 100000
    # CLI: only explicitly given flags override the preset
    import argparse  # noqa: F401
    eng, _ = hybrid._init_chains(HybridConfig.preset(1, chains=6), "cuda")
    assert eng.rng_kind == "philox" and eng.B == 6
    eng2, _ = hybrid._init_chains(HybridConfig.preset(1, chains=6, rng="pcg64"), "cuda")
    assert eng2.rng_kind == "pcg64" and eng2.pcg_state is not None


def test_mcmc_only_small():
    """BASELINE configs[0] shape: local-displacement MCMC only, per-chain well statistics from the device classifier
    cross-checked against the oracle on the final configurations."""
    from flowstate_b200.drivers import HybridConfig, run_mcmc_only
    from oracle import observables_ref as obr
    cfg = HybridConfig(particles=3, chains=12, equilibration_steps=400, adjusting_frequency=200, sampling_frequency=50,
                       production_steps=2000)
    out = run_mcmc_only(cfg, log=lambda *a: None)
    assert out["attempts"] == 12 * 2400 and 0 < out["accepted"] < out["attempts"]
    assert out["samples_per_chain"] == 40 and len(out["delta_f"]) == 12
    assert all(0.0 <= a <= 1.0 and 0.0 <= b <= 1.0 and a + b <= 1.0 + 1e-12 for a, b in zip(out["p_a"], out["p_b"]))
    eng = out["engine"]
    L = float(np.float32(np.sqrt(3 / 0.03)))
    pos = eng.pos.cpu().numpy()
    assert (pos >= 0).all() and (pos <= L).all()
    from flowstate_b200.drivers import observables
    cls, _, _ = observables._classify(pos, L / 2, cfg.r0)
    assert np.array_equal(cls.cpu().numpy(), obr.classify(pos, L / 2, cfg.r0))


def test_graphed_hybrid_round_equals_eager_round():
    """drivers/rounds.HybridRound: the whole round (forked sampling pass of the next proposals, local sweep, both
    log-densities, fused energy + acceptance) replayed from a CUDA graph against the same round issued eagerly - same
    Philox streams, same base noise (torch's generator re-seeded) -> identical accept masks, positions, energies and
    counters after every round, including the rounds that are replays."""
    import flowstate_b200.MCMC as MC
    import flowstate_b200.normflows as NF
    from flowstate_b200.drivers.rounds import HybridRound
    from oracle import energy_ref as er
    n, B, rho = 32, 256, 0.3
    L = float(np.float32(np.sqrt(n / rho)))
    bound = L / 2
    torch.manual_seed(0)
    flows = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, 2, 128, list(range(2 * n)), num_bins=8,
                                                             tail_bound=bound) for _ in range(3)]
    model = NF.NormalizingFlow(NF.Energy.UniformParticle(n, 2, bound, "cuda"), flows)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.02 * torch.randn_like(p))
    model = model.cuda().eval()
    pos, _ = er.batch_lattices(B, n, rho, seed0=3)
    out = {}
    for graph in (False, True):
        eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                                   initial_max_displacement=0.4, rng="philox", seeds=9)
        eng.set_nf_model(model)
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        hr = HybridRound(eng, model, 40, use_graph=graph)
        hist = []
        for r in range(6):
            mask = hr.step()
            hist.append((mask.clone(), eng.pos.clone(), eng.E.clone(), eng.accepted.clone()))
        out[graph] = hist
        if graph:
            assert hr.graphs[0] is not None and hr.graphs[1] is not None and hr.launches_per_round > 0
    for r, (a, b) in enumerate(zip(out[False], out[True])):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]), r
    # the chains moved (local acceptances) and every round attempted its global move
    assert int(out[True][-1][3].sum()) > 0 and not torch.equal(out[True][-1][1], out[True][0][1])
