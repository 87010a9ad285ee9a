"""GPU parity of the THROUGHPUT sweep kernel (local_sweep_fast_kernel: Philox streams, several chains per warp) -
the kernel bench.py and the drivers run.

(1) lock-step oracle: the kernel's per-step trace (particle index, e_old, e_new, decision) is checked step by step
    against oracle.mc_ref / energy_ref fed with the same Philox draws.  The oracle state follows the kernel's decision
    after every step, so EVERY decision of the run is checked, not only those before a first divergence: a decision
    may differ from the reference rule (MCMC/monte_carlo.py:191-223) only inside the epsilon band
    |log u - Delta| <= beta tol_E (M_old + M_new), tol_E = 1e-5 (M = sum of |pair terms| + |well term|,
    equal to |e| unless attractive and repulsive terms cancel - see oracle.mc_ref.lockstep_check); energies within
    1e-5 of M; the final positions must then be bit-equal and the counters equal.
(2) traced run == untraced run bit for bit (tracing must not change the trajectory).
(3) fast kernel vs the reference-order parity kernel (local_sweep_kernel) on the same Philox draws.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import energy_ref as er
from oracle import mc_ref as mr
from oracle import philox_ref as pr

POT = er.Potential(2, [-10.0, -10.5], 1.2, 15.0)
TOL_E = 1e-5
SEED = 0x1234ABCD5678


def _engine(pos, L, md, rng="philox", **kw):
    import flowstate_b200.MCMC as MC
    return MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, pos.shape[1], num_wells=2, V0_list=[-10.0, -10.5],
                                r0=1.2, k=15, initial_max_displacement=md, rng=rng, philox_seed=SEED, **kw)


def _start(n, rho, B, overlap):
    if n == 3:          # the reference's own N = 3 style: a small cluster inside the left well
        L = er.box_length(n, rho)
        rng = np.random.default_rng(5)
        base = np.array([[L / 4 - 0.7, L / 2], [L / 4 + 0.7, L / 2], [L / 4, L / 2 + 1.1]])
        pos = (base[None] + 0.05 * rng.standard_normal((B, n, 2))).astype(np.float32)
    else:
        pos, L = er.batch_lattices(B, n, rho, seed0=77)
    if overlap:          # particle 1 within the hard core of particle 0: the chain starts at E = inf
        pos[:, 1] = pos[:, 0] + np.float32(0.2)
    return pos, L


def _lockstep(pos0, L, md, chain_id, steps, acc, idx, e):
    """One chain through oracle.mc_ref.lockstep_check on the device's own Philox draws."""
    p_all, u_all = pr.step_draws(SEED, chain_id, 0, steps, len(pos0))
    r = mr.lockstep_check(pos0, L, md, POT, p_all, u_all, acc, idx, e, tol_e=TOL_E)
    assert r["outside_band"] == 0, r
    assert r["max_energy_err"] <= TOL_E, r
    return r["flips_in_band"], r["final"]


CASES = [  # n, rho, max_disp, steps, overlap start
    (3, 0.03, 0.65, 2500, False),
    (3, 0.03, 25.0, 2000, False),      # max_disp > L: the floor-mod wraps more than one box length
    (32, 0.03, 0.65, 2000, False),
    (32, 0.5, 0.4, 2000, False),
    (32, 0.5, 0.4, 2000, True),
    (64, 0.5, 0.4, 2000, False),
    (256, 0.5, 0.3, 2000, False),
    (256, 0.03, 0.65, 2000, False),
    (45, 0.4, 0.5, 2000, False),       # N not a multiple of the lane-group size (NaN padding of the last trip)
]


@pytest.mark.parametrize("n,rho,md,steps,overlap", CASES)
def test_fast_kernel_lockstep_oracle(n, rho, md, steps, overlap):
    B = 11                                    # not a multiple of the chains per warp: exercises the shadow groups
    pos, L = _start(n, rho, B, overlap)
    eng = _engine(pos, L, md, chain_id0=1000)
    half = steps // 3
    t1 = eng.particle_displacement(half, trace=True)       # split launch: step ids continue across launches
    t2 = eng.particle_displacement(steps - half, trace=True)
    acc = torch.cat([t1["accept"], t2["accept"]], 1).cpu().numpy()
    idx = torch.cat([t1["idx"], t2["idx"]], 1).cpu().numpy()
    e = torch.cat([t1["e"], t2["e"]], 1).cpu().numpy().astype(np.float64)
    # untraced run of the same chains: identical trajectory
    plain = _engine(pos, L, md, chain_id0=1000)
    plain.particle_displacement(steps)
    assert torch.equal(plain.pos, eng.pos) and torch.equal(plain.accepted, eng.accepted)
    assert torch.equal(plain.E.nan_to_num(nan=1e300), eng.E.nan_to_num(nan=1e300))
    assert torch.equal(plain.W.nan_to_num(nan=1e300), eng.W.nan_to_num(nan=1e300))
    total_flips = 0
    for c in (0, 4, B - 1):
        flips, state = _lockstep(pos[c], L, md, 1000 + c, steps, acc[c], idx[c], e[c])
        total_flips += flips
        np.testing.assert_array_equal(eng.pos[c].cpu().numpy(), state)
        assert int(eng.attempts[c].item()) == steps
        assert int(eng.accepted[c].item()) == int(acc[c].sum())
        Er, Wr = er.total_energy_virial(state.astype(np.float64), L, L, POT)
        if not overlap:
            assert abs(eng.E[c].item() - Er) <= TOL_E * max(1.0, abs(Er))
    frac = acc.mean()
    assert frac > 0.02, frac               # (dilute systems accept almost every move)
    print("n=%d rho=%g md=%g: %d decisions checked, %d inside the epsilon band differ, acceptance %.3f"
          % (n, rho, md, 3 * steps, total_flips, frac))


@pytest.mark.parametrize("n,rho,md", [(32, 0.5, 0.4), (256, 0.5, 0.3), (3, 0.03, 0.65)])
def test_fast_kernel_equals_reference_order_kernel(n, rho, md):
    """The same Philox draws through both kernels: decisions agree up to a first difference, which must sit inside
    the epsilon band of the slow kernel's own energies."""
    B, steps = 16, 1500
    pos, L = _start(n, rho, B, False)
    fast = _engine(pos, L, md)
    slow = _engine(pos, L, md, rng="philox_ref")
    tf = fast.particle_displacement(steps, trace=True)
    ts = slow.particle_displacement(steps, trace=True)
    af, a_s = tf["accept"].cpu().numpy(), ts["accept"].cpu().numpy()
    es = ts["e"].cpu().numpy().astype(np.float64)
    same = 0
    for c in range(B):
        diff = np.nonzero(af[c] != a_s[c])[0]
        first = diff[0] if len(diff) else steps
        np.testing.assert_array_equal(tf["idx"][c, :first + 1].cpu().numpy(), ts["idx"][c, :first + 1].cpu().numpy())
        ef = tf["e"][c, :first + 1].cpu().numpy().astype(np.float64)
        fin = np.isfinite(es[c, :first + 1])
        assert np.array_equal(np.isfinite(ef), fin)
        # the reference-order kernel forms the minimum image like the reference (subtract at magnitude L, then shift):
        # on close pairs that straddle the periodic boundary it is the less accurate of the two by up to ulp(L) / r
        assert (np.abs(ef[fin] - es[c, :first + 1][fin]) <= 2e-4 * np.maximum(1.0, np.abs(es[c, :first + 1][fin]))).all()
        if len(diff):
            eo, en = es[c, first]
            _, u = pr.step_draws(SEED, c, first, 1, n)
            eps = TOL_E * (abs(eo) + abs(en))
            assert abs(en - eo) <= eps or abs(np.log(u[0, 2]) + (en - eo)) <= eps, (c, first, eo, en, u[0, 2])
        else:
            same += 1
            assert torch.equal(fast.pos[c], slow.pos[c])
            assert int(fast.accepted[c]) == int(slow.accepted[c])
    assert same >= B // 2


def test_fast_kernel_lane_group_sizes_agree(monkeypatch):
    """8, 16 and 32 lanes per chain are the same computation up to the float32 summation order: the trajectories
    agree until a decision inside the epsilon band (rare), and most chains never see one."""
    import subprocess, sys, os, json
    code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from oracle import energy_ref as er
import flowstate_b200.MCMC as MC
pos, L = er.batch_lattices(24, 64, 0.5, seed0=5)
eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, 64, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                           initial_max_displacement=0.4, rng="philox", philox_seed=99)
eng.particle_displacement(1000)
print(json.dumps({"pos": eng.pos.cpu().numpy().tolist(), "acc": eng.accepted.cpu().numpy().tolist()}))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for lpc in ("8", "16", "32"):
        env = dict(os.environ, FS_SWEEP_LPC=lpc)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[lpc] = json.loads(r.stdout.strip().splitlines()[-1])
    ref = np.array(outs["8"]["pos"])
    for lpc in ("16", "32"):
        got = np.array(outs[lpc]["pos"])
        same = sum(int(np.array_equal(got[c], ref[c])) for c in range(len(ref)))
        assert same >= len(ref) - 4, (lpc, same)
