"""CPU tests of the host side: C-ABI surface, module tree / state_dict parity with the
reference, the train-mode autograd path against the golden vectors, fail-loud
behaviour without a GPU, RNG state mirroring, and the world_size-2 sharding logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import flowstate_b200
from flowstate_b200 import _lib
import flowstate_b200.normflows as NF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "flowstate_b200.h")).read()
    declared = set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().fs_version() >= 100


def test_cabi_rejects_bad_arguments_without_gpu():
    l = _lib.lib()
    pot = _lib.make_pot(2, [-10, -10.5], 1.2, 15)
    assert l.fs_energy_total(None, 1, 3, 10.0, 10.0, pot, None, None, None, None) == 1
    assert b"fs_energy_total" in l.fs_last_error()
    assert l.fs_local_sweep(None, None, None, None, None, None, 1, 3, 1, 10.0, 10.0, 1.0, pot, None, None, None,
                            None, None) == 1
    assert l.fs_flow_create(None, None) == 1


def _sd(g, prefix="sd__"):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _build(n, K, blocks, H, nb, bound):
    base = NF.Energy.UniformParticle(n, 2, bound)
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, blocks, H, range(2 * n), num_bins=nb,
                                                              tail_bound=bound) for _ in range(K)]
    return NF.NormalizingFlow(base, layers)


def test_state_dict_matches_reference_names_and_seeded_init(golden_dir):
    g = np.load(os.path.join(golden_dir, "flow_init_seed123.npz"))
    ref = _sd(g)
    torch.manual_seed(123)
    model = _build(3, 2, 2, 16, 8, 5.0)
    mine = model.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].shape == ref[k].shape and mine[k].dtype == ref[k].dtype, k
        assert torch.equal(mine[k], ref[k]), "seeded initialisation differs at %s" % k


@pytest.mark.parametrize("tag", ["n3_k3", "n4_k4", "n32_k2"])
def test_autograd_path_matches_reference(golden_dir, tag):
    """Train-mode arithmetic (with BatchNorm modules left in eval so statistics match)."""
    g = np.load(os.path.join(golden_dir, "flow_%s.npz" % tag))
    model = _build(int(g["n"]), int(g["K"]), int(g["blocks"]), int(g["H"]), int(g["nb"]), float(g["bound"]))
    model.load_state_dict(_sd(g))
    model.eval()
    for f in model.flows:
        f.training = True            # coupling takes the autograd path, BN keeps running stats
    model.training = True
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        z, ld = model.inverse_and_log_det(x)
        xf, ldf = model.forward_and_log_det(torch.from_numpy(g["z0"]))
    np.testing.assert_allclose(z.numpy(), g["inv_z"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ld.numpy(), g["inv_ld"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(xf.numpy(), g["fwd_x"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ldf.numpy(), g["fwd_ld"], rtol=1e-4, atol=1e-4)
    # gradients flow to every learnable tensor that the reference trains
    model.zero_grad()
    loss = model.forward_kld(x[3:])
    loss.backward()
    missing = [n for n, p in model.named_parameters() if p.grad is None and "preprocessing.weights" not in n]
    assert not missing, missing


def test_eval_mode_fails_loudly_without_cuda():
    model = _build(3, 2, 2, 16, 8, 5.0).eval()
    x = torch.zeros(4, 6)
    with pytest.raises(flowstate_b200.FlowStateError):
        model.log_prob(x)
    with pytest.raises(flowstate_b200.FlowStateError):
        model.sample(2)
    if not torch.cuda.is_available():
        import flowstate_b200.MCMC as MC
        with pytest.raises(flowstate_b200.FlowStateError):
            MC.BatchedMonteCarlo(np.zeros((2, 3, 2), np.float32), MC.SimulationBox(10.0), 1.0, 3)
        with pytest.raises(flowstate_b200.FlowStateError):
            MC.SimulationBox(10.0).apply_pbc(np.array([1.0, 2.0]))


def test_value_errors_match_reference_conventions():
    layer = NF.flows.CircularCoupledRationalQuadraticSpline(6, 2, 16, range(6), num_bins=8, tail_bound=5.0)
    with pytest.raises(ValueError):
        layer.inverse(torch.zeros(6))
    with pytest.raises(ValueError):
        layer.forward(torch.zeros(2, 5))


def test_pcg64_state_roundtrip():
    from flowstate_b200.MCMC.batched import pcg64_set_state, pcg64_state_words
    a = np.random.default_rng(42)
    a.integers(7)                      # leaves a buffered 32-bit half behind
    w = pcg64_state_words(a)
    assert w[4] == 1
    b = np.random.default_rng(0)
    pcg64_set_state(b, w)
    assert [a.integers(7), a.random(), a.integers(1000)] == [b.integers(7), b.random(), b.integers(1000)]


def test_shard_range_covers_all_chains():
    from flowstate_b200.parallel import shard_range
    for total in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    from flowstate_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)              # different weights per rank before the broadcast
        model = _build(3, 2, 2, 16, 8, 5.0)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.01 * torch.randn_like(p))
        parallel.broadcast_flow(model, src=0)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same_weights = all(torch.equal(gathered[0], t) for t in gathered)
        # data-parallel step: each rank a different half of one batch; averaged gradients must
        # equal the single-process gradient of the mean loss over the whole batch (BN in eval)
        model.eval()
        for f in model.flows:
            f.training = True
        model.training = True
        g = torch.Generator().manual_seed(7)
        x = (torch.rand(8, 6, generator=g) * 2 - 1) * 5
        model.zero_grad()
        model.forward_kld(x).backward()
        full = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()
        model.zero_grad()
        s, c = parallel.shard_range(8, rank, world)
        model.forward_kld(x[s:s + c]).backward()
        parallel.allreduce_gradients(model)
        mine = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        att, acc = parallel.allreduce_counters(torch.tensor([3 + rank]), torch.tensor([1 + rank]))
        ret[rank] = (same_weights, float((mine - full).abs().max()), float(full.abs().max()), att, acc)
    finally:
        dist.destroy_process_group()


def test_world_size_2_broadcast_and_gradient_allreduce():
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    for r in range(world):
        same, err, scale, att, acc = res[r]
        assert same
        assert err <= 1e-5 * max(1.0, scale), (err, scale)
        assert att == 7 and acc == 3


def _trainer_worker(rank, world, port, ret):
    import torch.distributed as dist
    from flowstate_b200 import parallel
    from flowstate_b200.drivers import hybrid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(5)                                   # same initial weights, as _build_flow broadcasts them
        model = _build(3, 2, 2, 16, 8, 5.0)
        cfg = hybrid.HybridConfig.preset(2, particles=3, K=2, blocks=2, hidden=16, bins=8, batch_size=8, epochs=2,
                                         lr=1e-2, cuda_graph=0)
        g = torch.Generator().manual_seed(40 + rank)
        rows = 30 if rank == 0 else 19                         # unequal data: 4 vs 3 minibatches -> 3 on both ranks
        data = (torch.rand(rows, 6, generator=g) * 2 - 1) * 5
        if rank == 1:
            data[8:16] = float("nan")                          # a non-finite loss on ONE rank: both must skip that step
        torch.manual_seed(9)                                   # same permutations on both ranks (keeps the NaN batch aligned)
        losses = hybrid._train(model, data, cfg, cfg.epochs)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        ret[rank] = (all(torch.equal(gathered[0], t) for t in gathered), bool(torch.isfinite(flat).all()),
                     [float(x) for x in losses], model.training)
    finally:
        dist.destroy_process_group()


def test_world_size_2_trainer_keeps_collectives_aligned():
    """drivers.training.FlowTrainer through hybrid._train on two gloo ranks with unequal sample counts and a NaN batch on
    one rank: no deadlock, identical finite weights afterwards (rank-invariant collective count, collective
    skip decision)."""
    import torch.multiprocessing as mp
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_trainer_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    for r in range(world):
        same, finite, losses, training = res[r]
        assert same and finite and not training
        assert len(losses) == 2 and all(np.isfinite(l) for l in losses)


def test_drop_in_top_level_import_names():
    """The reference drivers do `import normflows as NF; import MCMC as MC`
    (hybrid_NF_MCMC/main_algorithm_1.py:29-30): both names must resolve to this package when its
    directory is placed on sys.path."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import normflows as NF, MCMC as MC;"
            "assert 'flowstate_b200' in NF.__file__ and 'flowstate_b200' in MC.__file__;"
            "NF.flows.CircularCoupledRationalQuadraticSpline; NF.Energy.UniformParticle; NF.NormalizingFlow;"
            "MC.MonteCarlo; MC.EnergyCalculator; MC.SimulationBox; MC.initialise_low_left; print('ok')"
            % (ROOT, os.path.join(ROOT, "flowstate_b200")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_observables_host_statistics_match_reference_golden(golden_dir, monkeypatch):
    """The host half of flowstate_b200.drivers.observables (cumulative p_A / p_B / dF, g(r) normalisation) against the
    reference's golden outputs, with the device kernels replaced by the oracle (no GPU here)."""
    import torch
    from flowstate_b200.drivers import observables as obs
    from oracle import observables_ref as obr
    g = np.load(os.path.join(golden_dir, "observables.npz"))

    def fake_classify(configurations, half_box, r0):
        cfg = np.asarray(configurations)
        cls = obr.classify(cfg, half_box, r0)
        state = np.where((cls == 1).all(axis=1), 1, np.where((cls == 2).all(axis=1), 2, 0)).astype(np.uint8)
        return torch.from_numpy(cls), torch.from_numpy(state), torch.from_numpy(cfg[:, :, 0].astype(np.float64).mean(axis=1))

    def fake_hist(final_samples, bound, dr):
        return torch.from_numpy(obr.pair_histogram(np.asarray(final_samples), bound, dr).astype(np.int32))

    monkeypatch.setattr(obs, "_classify", fake_classify)
    monkeypatch.setattr(obs, "pair_histogram", fake_hist)
    half_box, r0, start = float(g["ws_half_box"]), float(g["ws_r0"]), int(g["ws_start"])
    avg_x, p_a, p_b, dF, runs = obs.calculate_well_statistics(g["ws_cfgs"], start, half_box, r0)
    np.testing.assert_allclose(avg_x, g["ws_avg_x"], rtol=1e-6)
    assert np.array_equal(p_a, g["ws_p_a"]) and np.array_equal(p_b, g["ws_p_b"])
    np.testing.assert_allclose(dF, g["ws_dF"], rtol=0, atol=1e-15)
    names = obs.classify_particles(g["ws_cfgs"], half_box, r0)
    assert names.shape == g["ws_class"].shape and set(np.unique(names)) <= {"A", "B", "Outside"}
    for tag in "abc":
        r, gr = obs.calculate_pair_correlation(g["pc_%s_samples" % tag], int(g["pc_%s_n" % tag]),
                                               float(g["pc_%s_bound" % tag]), float(g["pc_%s_dr" % tag]))
        assert np.array_equal(r, g["pc_%s_r" % tag])
        np.testing.assert_allclose(np.asarray(gr), g["pc_%s_g" % tag], rtol=1e-14, atol=0)


def test_driver_local_phase_schedule():
    """_local_phase splits a run of local moves at the reference's adaptation / sampling boundaries
    (main_algorithm_1.py:203-232: adjust every ADJUSTING_FREQUENCY steps, sample every SAMPLING_FREQUENCY steps,
    both counted from the chain's first step)."""
    import torch
    from flowstate_b200.drivers import hybrid

    class FakeEngine:
        def __init__(self):
            self.steps, self.adjusted_at, self.pos = 0, [], torch.zeros(2, 3, 2)

        def particle_displacement(self, n):
            assert n >= 1
            self.steps += n

        def adjust_displacement(self):
            self.adjusted_at.append(self.steps)

        def centred(self, pos):
            return torch.full((2, 6), float(self.steps))

    cfg = hybrid.HybridConfig(adjusting_frequency=300, sampling_frequency=70)
    eng, samples = FakeEngine(), []
    end = hybrid._local_phase(eng, 1000, cfg, collect=samples, step0=0)
    assert end == 1000 and eng.steps == 1000
    assert eng.adjusted_at == [300, 600, 900]
    assert [int(t[0, 0]) for t in samples] == list(range(70, 1001, 70))
    # continuing from a non-zero step count keeps the global phase of both schedules
    end = hybrid._local_phase(eng, 500, cfg, collect=samples, step0=end)
    assert end == 1500 and eng.adjusted_at == [300, 600, 900, 1200, 1500]
    assert [int(t[0, 0]) for t in samples][-7:] == [1050, 1120, 1190, 1260, 1330, 1400, 1470]
    # without collection only the adaptation boundaries split the run
    eng2 = FakeEngine()
    hybrid._local_phase(eng2, 650, hybrid.HybridConfig(adjusting_frequency=0, sampling_frequency=70))
    assert eng2.steps == 650 and eng2.adjusted_at == []


def test_bench_helpers_without_gpu():
    """bench.py's algorithmic-flop formula reproduces SURVEY 8(a23)'s per-sample figures, and the clock sampler degrades
    to an empty record on a machine without NVML / nvidia-smi instead of failing the run."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fs_bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = bench.WORKLOADS
    assert abs(bench.flops_per_sample_layer(w["alg1_n32"]) / 1e6 - 10.0) < 0.05       # C2: 10.0 MF
    assert abs(bench.flops_per_sample_layer(w["alg1_n256"]) / 1e6 - 21.4) < 0.1       # C3: 21.4 MF
    assert abs(bench.flops_per_sample_layer(w["alg2_n64"]) / 1e6 - 0.92) < 0.01       # C4: 0.92 MF
    cs = bench.ClockSampler(0)
    cs.wait_ready(timeout=0.2)
    out = cs.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"} and isinstance(out["reasons"], list)


def test_initialisers_match_reference_golden(golden_dir):
    """MCMC/initialise.py:8-305 restated: every low-left / low-right grid for N = 1..12 and the FCC-like lattices,
    bit-identical to arrays produced by the reference's own functions (oracle/make_golden.py: gen_initialise)."""
    import flowstate_b200.MCMC as MC
    g = np.load(os.path.join(golden_dir, "initialise.npz"))
    checked = 0
    for k in g.files:
        if k.endswith("_box"):
            continue
        parts = k.split("_")
        n, rho = int(parts[1][1:]), float(parts[2][3:])
        if k.startswith("low"):
            fn = MC.initialise_low_left if k[3] == "L" else MC.initialise_low_right
            p, box = fn(n, rho, 2.0 if k.endswith("ar2") else 1.0)
        else:
            p, box = MC.initialise_fcc(n, rho, float(parts[3][2:]))
        assert np.array_equal(p, g[k]), k
        assert np.array_equal([box.box_size_x, box.box_size_y], g[k + "_box"]), k
        checked += 1
    assert checked == 53
    with pytest.raises(ValueError):
        MC.initialise_low_left(13, 0.03)
    pos, box = MC.initialise_chains(5, 3, 0.03, first_chain=2)
    left, _ = MC.initialise_low_left(3, 0.03)
    right, _ = MC.initialise_low_right(3, 0.03)
    assert pos.shape == (5, 3, 2) and pos.dtype == np.float32
    assert np.array_equal(pos[0], left.astype(np.float32)) and np.array_equal(pos[1], right.astype(np.float32))
    assert np.array_equal(pos[2], pos[0]) and abs(box.box_size_x - 10.0) < 1e-12


def test_trainer_optimizer_reset_equals_a_new_adam():
    """FlowTrainer.fresh_optimizer zeroes the existing Adam state in place from the second cycle on; the updates that
    follow must be those of a newly built optimizer (main_algorithm_2.py:440 builds one per cycle)."""
    import flowstate_b200.normflows as NF
    from flowstate_b200.drivers.training import FlowTrainer
    n, bound = 4, 3.0
    xs = [(torch.rand(16, 2 * n, generator=torch.Generator().manual_seed(i)) * 2 - 1) * bound for i in range(4)]
    out = []
    for reuse in (True, False):
        torch.manual_seed(0)
        base = NF.Energy.UniformParticle(n, 2, bound)
        model = NF.NormalizingFlow(base, [NF.flows.CircularCoupledRationalQuadraticSpline(
            2 * n, 1, 16, range(2 * n), num_bins=4, tail_bound=bound) for _ in range(2)]).train()
        tr = FlowTrainer(model, 1e-2, 1e-4, 1.0, 16, use_graph=False)
        tr.fresh_optimizer()
        tr.step(xs[0]); tr.step(xs[1])
        if reuse:
            first = tr.opt
            assert tr.fresh_optimizer() is first              # same object, state zeroed
        else:
            tr.opt = None
            tr.fresh_optimizer()
        tr.step(xs[2]); tr.step(xs[3])
        out.append(torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    assert torch.equal(out[0], out[1])


def test_trainer_flat_bucket_layout():
    """The gradient bucket of FlowTrainer: every trainable parameter's .grad is a view into ONE flat buffer, tensors
    start on 16-byte boundaries (the training kernels load float4 through these pointers), the padding stays zero and
    parameters without a gradient (PeriodicFeaturesElementwise.weights) are left alone."""
    import flowstate_b200.normflows as NF
    from flowstate_b200.drivers.training import FlowTrainer
    n, bound = 3, 3.0                                       # N = 3, nb = 5: tensor sizes that are not multiples of 4
    torch.manual_seed(0)
    base = NF.Energy.UniformParticle(n, 2, bound)
    model = NF.NormalizingFlow(base, [NF.flows.CircularCoupledRationalQuadraticSpline(
        2 * n, 1, 10, range(2 * n), num_bins=5, tail_bound=bound) for _ in range(2)]).train()
    tr = FlowTrainer(model, 1e-2, 0.0, 1.0, 8, use_graph=False)
    x = (torch.rand(8, 2 * n) * 2 - 1) * bound
    assert tr.step(x) is not None
    assert not tr.native_adam and tr.flat_p is None         # CPU: torch.optim.Adam on the parameters where they are
    lo, hi = tr.flat.data_ptr(), tr.flat.data_ptr() + tr.flat.numel() * 4
    covered = torch.zeros(tr.flat.numel(), dtype=torch.bool)
    for p in tr.trainable:
        assert p.grad is not None and lo <= p.grad.data_ptr() < hi and (p.grad.data_ptr() - lo) % 16 == 0
        off = (p.grad.data_ptr() - lo) // 4
        assert not covered[off:off + p.numel()].any()
        covered[off:off + p.numel()] = True
    assert tr.flat.numel() % 4 == 0 and (tr.flat[~covered] == 0).all() and tr.flat[covered].abs().sum() > 0
    skipped = [k for k, p in model.named_parameters() if p.grad is None]
    assert skipped and all(k.endswith("preprocessing.weights") for k in skipped), skipped
