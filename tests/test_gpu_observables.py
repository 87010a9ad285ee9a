"""Observables through the C ABI against the golden outputs of the reference's
hybrid_NF_MCMC/utils.py functions and against the oracle on larger seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import observables_ref as obr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "observables.npz")


def _obs():
    from flowstate_b200.drivers import observables
    return observables


def test_well_statistics_match_reference_golden():
    g = np.load(GOLD)
    obs = _obs()
    half_box, r0, start = float(g["ws_half_box"]), float(g["ws_r0"]), int(g["ws_start"])
    cls = obs.classify_particles(g["ws_cfgs"], half_box, r0)
    code = np.zeros(cls.shape, dtype=np.uint8)
    code[cls == "A"] = 1
    code[cls == "B"] = 2
    assert np.array_equal(code, g["ws_class"])                       # bit-exact classification
    avg_x, p_a, p_b, dF, runs = obs.calculate_well_statistics(g["ws_cfgs"], start, half_box, r0)
    np.testing.assert_allclose(avg_x, g["ws_avg_x"], rtol=1e-6)      # np.mean of float32 accumulates in float32
    assert np.array_equal(p_a, g["ws_p_a"]) and np.array_equal(p_b, g["ws_p_b"])
    np.testing.assert_allclose(dF, g["ws_dF"], rtol=0, atol=1e-15)
    assert list(runs) == list(g["ws_runs"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_pair_correlation_matches_reference_golden(tag):
    g = np.load(GOLD)
    obs = _obs()
    samples, n, bound, dr = g["pc_%s_samples" % tag], int(g["pc_%s_n" % tag]), float(g["pc_%s_bound" % tag]), float(g["pc_%s_dr" % tag])
    counts = obs.pair_histogram(samples, bound, dr).cpu().numpy()
    assert np.array_equal(counts, obr.pair_histogram(samples, bound, dr))    # integer histogram: bit-exact
    r, gr = obs.calculate_pair_correlation(samples, n, bound, dr)
    assert np.array_equal(r, g["pc_%s_r" % tag])
    np.testing.assert_allclose(np.asarray(gr), g["pc_%s_g" % tag], rtol=1e-14, atol=0)


def test_observables_at_bench_sizes_against_oracle():
    rng = np.random.default_rng(5)
    obs = _obs()
    for n, B in ((32, 64), (256, 6)):
        bound = float(np.float32(np.sqrt(n / 0.03))) / 2
        x = rng.uniform(-bound, bound, size=(B, n, 2)).astype(np.float32)
        dr = bound / 49.5
        got = obs.pair_histogram(x, bound, dr).cpu().numpy()
        assert np.array_equal(got, obr.pair_histogram(x, bound, dr))
        assert got.sum() <= B * n * (n - 1)
        pos = (x + np.float32(bound)).astype(np.float32)
        cls, state, avg = obs._classify(pos, bound, 1.2)
        assert np.array_equal(cls.cpu().numpy(), obr.classify(pos, bound, 1.2))
        np.testing.assert_allclose(avg.cpu().numpy(), pos[:, :, 0].astype(np.float64).mean(axis=1), rtol=1e-12)
        assert int(state.sum().item()) == 0                              # scattered particles: never all in one well


def test_run_outputs_formats(tmp_path):
    obs = _obs()
    samples = [(i, -1.5 + i, 0.03, 0.02, 10.0, 10.0, np.arange(6, dtype=np.float32).reshape(3, 2) + i) for i in range(4)]
    obs.save_run_outputs(str(tmp_path), samples, testing_samples=[s[6] for s in samples[:2]])
    rows = open(tmp_path / "sampled_data.csv").read().splitlines()
    assert rows[0] == "cycle_number,energy_per_particle,density,pressure,box_size_x,box_size_y,particle_configuration"
    assert len(rows) == 5 and rows[1].startswith("0,-1.5,0.03,0.02,10.0,10.0,")
    assert np.load(tmp_path / "mc_run_configs.npy").shape == (4, 3, 2)
    assert np.load(tmp_path / "mc_run_testing_configs.npy").shape == (2, 3, 2)
