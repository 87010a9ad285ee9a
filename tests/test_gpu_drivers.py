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
    cfg = HybridConfig(particles=4, chains=32, equilibration_steps=300, adjusting_frequency=10000, K=2, blocks=2,
                       hidden=128, bins=8, lr=1e-3, batch_size=64, cycles=5, local_steps=40)
    out = run_algorithm_2(cfg, log=lambda *a: None)
    assert out["attempts"] == 32 * (300 + 5 * 41)
    assert np.isfinite(out["final_loss"])


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
