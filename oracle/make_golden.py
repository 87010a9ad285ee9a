"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # needs /root/reference

Every fixture holds seeded, float32-representable inputs together with what the
reference's own code returned for them, so the oracle (and through it the CUDA
path) can be pinned on machines where the reference tree does not exist.
Reference entry points exercised:
  MCMC/potential.py:3-29, 55-116; MCMC/simulation_box.py:19-65;
  MCMC/energy_calculator.py:48-203; MCMC/monte_carlo.py:146-303, 375-403;
  NF/normflows/core.py:178-214 with flows built exactly like
  hybrid_NF_MCMC/main_algorithm_1.py:277-284.
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import _refimport  # noqa: E402
from oracle import energy_ref as er  # noqa: E402
from oracle.mc_ref import SpyRNG  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
POT = dict(num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15)


def quiet_logger():
    lg = logging.getLogger("fs_golden")
    lg.setLevel(logging.CRITICAL)
    lg.addHandler(logging.NullHandler())
    return lg


# ---------------------------------------------------------------------------
def gen_energy(ref):
    SB = ref["simulation_box"].SimulationBox
    EC = ref["energy_calculator"].EnergyCalculator
    pot = ref["potential"]
    out = {}
    # Known-answer inputs of SURVEY.md Appendix B
    r = np.array([0.5, 1.0, 2 ** (1 / 6), 2.0, 2.5, 2.5000001, 3.0])
    e, w = pot.lennard_jones_energy_virial(r)
    out["kat_lj_r"], out["kat_lj_e"], out["kat_lj_w"] = r, e, w
    P = np.array([[2.5, 5], [7.5, 5], [5, 5], [0, 0], [3.7, 5], [9.9, 5]], dtype=np.float64)
    out["kat_dw_pos"] = P
    out["kat_dw_v"] = pot.double_well_potential(P, 10, 10, [-10, -10.5], 1.2, 15, 2)
    box = SB(10.0)
    out["kat_pbc_in"] = np.array([[-0.1, 10.0], [10.3, -1e-17]])
    out["kat_pbc_out"] = np.array([box.apply_pbc(p) for p in out["kat_pbc_in"]])

    cases = []
    # (name, positions float32-exact, L, wells on?)
    cases.append(("kat3", np.array([[2.0, 5.0], [3.5, 5.0], [2.75, 6.3]], np.float32), 10.0, True))
    cases.append(("kat3_nowell", np.array([[0.2, 0.3], [9.7, 9.9], [5, 5]], np.float32), 10.0, False))
    cases.append(("kat3_overlap", np.array([[1, 1], [1.3, 1], [5, 5]], np.float32), 10.0, True))
    for n in (3, 32, 64, 256):
        for rho in (0.03, 0.5):
            for seed in (0, 1):
                p, L = er.jittered_lattice(n, rho, 1000 * n + seed)
                cases.append(("lat_n%d_rho%g_s%d" % (n, rho, seed), p, L, True))
    # denser, cancellation-heavy and a clustered init (spacing 1.5 around a well)
    p, L = er.jittered_lattice(64, 0.8, 7, jitter=0.3)
    cases.append(("dense_n64", p, L, True))
    L = er.box_length(32, 0.03)
    g = np.array([(i, j) for i in range(6) for j in range(6)][:32], dtype=np.float64)
    clus = (g - g.mean(0)) * 1.5 + np.array([L / 4, L / 2])
    clus += (np.random.default_rng(5).random((32, 2)) - 0.5) * 0.2
    cases.append(("cluster_n32", clus.astype(np.float32), L, True))
    # one particle copied to within 0.3 of another -> inf
    p, L = er.jittered_lattice(32, 0.5, 99)
    p = p.copy()
    p[5] = p[17] + np.float32(0.2)
    cases.append(("overlap_n32", p, L, True))
    # pair straddling the periodic boundary
    L = 12.0
    p = np.array([[0.1, 6.0], [11.6, 6.2], [6.0, 0.2], [6.3, 11.7], [3.0, 6.0]], np.float32)
    cases.append(("wrap_n5", p, L, True))

    names = []
    for name, p32, L, wells in cases:
        n = len(p32)
        kw = dict(POT) if wells else dict(num_wells=0, V0_list=[0, 0], r0=1.2, k=15)
        for mode, arr in (("f64", p32.astype(np.float64)), ("f32", p32.copy())):
            ec = EC(n, arr, SB(L), timing=False, **kw)
            E, W = ec.total_energy, ec.total_virial
            idx = sorted(set([0, n // 2, n - 1]))
            pe = np.array([ec.calculate_particle_energy_virial(arr, i) for i in idx], dtype=np.float64)
            out["%s__%s_E" % (name, mode)] = np.float64(E)
            out["%s__%s_W" % (name, mode)] = np.float64(W)
            out["%s__%s_pe" % (name, mode)] = pe
        out[name + "__pos"] = p32
        out[name + "__L"] = np.float64(L)
        out[name + "__wells"] = np.int64(1 if wells else 0)
        out[name + "__pidx"] = np.array(idx)
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "energy_cases.npz"), **out)
    print("energy_cases: %d cases" % len(names))


# ---------------------------------------------------------------------------
def gen_mc(ref):
    SB = ref["simulation_box"].SimulationBox
    MC = ref["monte_carlo"].MonteCarlo
    lg = quiet_logger()
    out = {}
    names = []
    specs = [("n3", 3, 0.03, 400, 0.65), ("n32", 32, 0.5, 300, 0.65), ("n64", 64, 0.5, 200, 0.4),
             ("n256", 256, 0.5, 60, 0.3)]
    for tag, n, rho, steps, md in specs:
        for mode in ("f32", "f64"):
            for seed in (42, 43):
                if n == 3:
                    # clustered start near the left / right well like initialise_low_*
                    L = er.box_length(n, rho)
                    cx = L / 4 if seed % 2 == 0 else 3 * L / 4
                    p32 = (np.array([[cx - 0.8, L / 2], [cx + 0.8, L / 2], [cx, L / 2 + 1.3]])).astype(np.float32)
                else:
                    p32, L = er.jittered_lattice(n, rho, seed)
                arr = p32.astype(np.float64) if mode == "f64" else p32.copy()
                mc = MC(arr, SB(L), 1.0, n, initial_max_displacement=md, target_acceptance=0.5,
                        timing=False, checking=False, logger=lg, seed=seed, **POT)
                spy = SpyRNG(mc.rng)
                mc.rng = spy
                E0 = mc.energy_calculator.total_energy
                # wrap particle energy to record (eno, enn) per step
                rec = []
                orig = mc.energy_calculator.calculate_particle_energy_virial

                def wrapped(pos, i, _o=orig, _r=rec):
                    v = _o(pos, i)
                    _r.append(v)
                    return v
                mc.energy_calculator.calculate_particle_energy_virial = wrapped
                acc = np.zeros(steps, np.uint8)
                half = steps // 2
                md_mid = None
                for s in range(steps):
                    a0 = mc.accepted_displacement
                    mc.particle_displacement()
                    acc[s] = mc.accepted_displacement - a0
                    if s + 1 == half:
                        mc.adjust_displacement()
                        md_mid = mc.max_displacement
                mc.adjust_displacement()
                key = "%s_%s_s%d" % (tag, mode, seed)
                names.append(key)
                ee = np.array(rec, dtype=np.float64).reshape(steps, 2, 2)
                out[key + "__pos0"] = p32
                out[key + "__L"] = np.float64(L)
                out[key + "__seed"] = np.int64(seed)
                out[key + "__steps"] = np.int64(steps)
                out[key + "__md0"] = np.float64(md)
                out[key + "__E0"] = np.float64(E0)
                out[key + "__idx"] = np.array(spy.ints, np.int64)
                out[key + "__u"] = np.array(spy.uniforms, np.float64)
                out[key + "__eno"] = ee[:, 0, 0]
                out[key + "__enn"] = ee[:, 1, 0]
                out[key + "__viro"] = ee[:, 0, 1]
                out[key + "__virn"] = ee[:, 1, 1]
                out[key + "__acc"] = acc
                out[key + "__posF"] = np.asarray(mc.particles, dtype=np.float64)
                out[key + "__EF"] = np.float64(mc.energy_calculator.total_energy)
                out[key + "__WF"] = np.float64(mc.energy_calculator.total_virial)
                out[key + "__md_mid"] = np.float64(md_mid)
                out[key + "__mdF"] = np.float64(mc.max_displacement)
                out[key + "__attempts"] = np.int64(mc.attempts_displacement)
                out[key + "__accepted"] = np.int64(mc.accepted_displacement)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "mc_local.npz"), **out)
    print("mc_local: %d traces" % len(names))


# ---------------------------------------------------------------------------
def build_ref_flow(ref, n, K, blocks, H, nb, bound, seed, sigma):
    """Constructor call of main_algorithm_1.py:277-284, then the SURVEY 8(d)
    perturbation so the flow is not the identity."""
    NF = ref["normflows"]
    torch.manual_seed(seed)
    with ref["quiet"]():
        base = NF.Energy.UniformParticle(n, 2, bound, device="cpu")
        layers = [NF.flows.CircularCoupledRationalQuadraticSpline(
            n * 2, blocks, H, range(n * 2), num_bins=nb, tail_bound=bound) for _ in range(K)]
        model = NF.NormalizingFlow(base, layers)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            prm.add_(sigma * torch.randn(prm.shape, generator=g))
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    model.eval()
    return model


FLOW_SPECS = [
    # tag, n, K, blocks, H, nb, rho, sigma, B
    ("n3_k3", 3, 3, 2, 32, 8, 0.03, 0.05, 64),        # odd N: couplings alternate (A.4-Q2)
    ("n4_k4", 4, 4, 2, 32, 15, 0.03, 0.05, 64),       # even N: x only element-wise
    ("n32_k2", 32, 2, 3, 64, 32, 0.03, 0.02, 48),     # Alg-1 bins (32) at N=32, cut-down width
    ("n4_k23", 4, 23, 2, 32, 15, 0.03, 0.05, 32),     # Alg-2 depth and bins (K=23, nb=15), cut-down width
    # hidden widths of the two driver architectures, so reference-generated vectors reach the tensor-core path
    ("n8_h128", 8, 3, 2, 128, 15, 0.03, 0.05, 40),    # Alg-2 width / blocks / bins (main_algorithm_2.py:62-70)
    ("n6_h256", 6, 2, 1, 256, 32, 0.03, 0.03, 40),    # Alg-1 width / bins (main_algorithm_1.py:63-70), one block
]


def gen_flow(ref, only=None):
    for tag, n, K, blocks, H, nb, rho, sigma, B in FLOW_SPECS:
        if only and tag not in only:
            continue
        bound = er.box_length(n, rho) / 2
        model = build_ref_flow(ref, n, K, blocks, H, nb, bound, seed=11, sigma=sigma)
        g = torch.Generator().manual_seed(3)
        x = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
        x[0, 0] = bound            # exactly on the upper edge: last bin (A.4-Q13)
        x[1, 1] = -bound
        x[2, 3] = bound * 1.25     # outside: passes through, base log-prob -inf
        z0 = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
        with torch.no_grad():
            lp = model.log_prob(x.clone())
            zi, ldi = model.inverse_and_log_det(x.clone())
            xf, ldf = model.forward_and_log_det(z0.clone())
            first, ld_first = model.flows[K - 1].inverse(x.clone())
        out = {"sd__" + k: v.numpy() for k, v in model.state_dict().items()}
        out.update(n=np.int64(n), K=np.int64(K), blocks=np.int64(blocks), H=np.int64(H), nb=np.int64(nb),
                   bound=np.float64(bound), x=x.numpy(), z0=z0.numpy(), log_prob=lp.numpy(),
                   inv_z=zi.numpy(), inv_ld=ldi.numpy(), fwd_x=xf.numpy(), fwd_ld=ldf.numpy(),
                   lastlayer_inv=first.numpy(), lastlayer_ld=ld_first.numpy())
        np.savez_compressed(os.path.join(GOLD, "flow_%s.npz" % tag), **out)
        print("flow_%s: params %d" % (tag, sum(p.numel() for p in model.parameters())))


# ---------------------------------------------------------------------------
def gen_global(ref):
    """nf_big_move traces (monte_carlo.py:235-303) with a small perturbed flow."""
    SB = ref["simulation_box"].SimulationBox
    MC = ref["monte_carlo"].MonteCarlo
    lg = quiet_logger()
    out = {}
    names = []
    for tag, n, rho, K, blocks, H, nb, sigma in (("n3", 3, 0.03, 3, 2, 32, 8, 0.05),
                                                 ("n8", 8, 0.03, 4, 2, 32, 15, 0.05)):
        L = er.box_length(n, rho)
        bound = L / 2
        model = build_ref_flow(ref, n, K, blocks, H, nb, bound, seed=21, sigma=sigma)
        for k, v in model.state_dict().items():
            out["%s__sd__%s" % (tag, k)] = v.numpy()
        out[tag + "__bound"] = np.float64(bound)
        out[tag + "__L"] = np.float64(L)
        torch.manual_seed(5)
        for seed in (42, 43, 44):
            if n == 3:
                cx = L / 4 if seed % 2 == 0 else 3 * L / 4
                p32 = (np.array([[cx - 0.8, L / 2], [cx + 0.8, L / 2], [cx, L / 2 + 1.3]])).astype(np.float32)
            else:
                p32, _ = er.jittered_lattice(n, rho, seed)
            mc = MC(p32.astype(np.float64), SB(L), 1.0, n, initial_max_displacement=0.65,
                    timing=False, checking=False, logger=lg, seed=seed, device=torch.device("cpu"), **POT)
            mc.set_nf_model(model)
            spy = SpyRNG(mc.rng)
            mc.rng = spy
            rounds, local = 12, 25
            with torch.no_grad():
                props = model.sample(rounds).reshape(rounds, n, 2).numpy() + bound   # main_algorithm_2.py:479-482
            # make some proposals near the current state so a few get accepted
            rec = dict(acc=[], eno=[], enn=[], pos_before=[], E_after=[])
            for r in range(rounds):
                for _ in range(local):
                    mc.particle_displacement()
                cfg = props[r].copy()
                if r % 3 == 2:
                    cfg = (np.asarray(mc.particles, dtype=np.float32)
                           + np.float32(0.01) * (r + 1)).astype(np.float32) % np.float32(L)
                    props[r] = cfg
                rec["pos_before"].append(np.asarray(mc.particles, dtype=np.float64).copy())
                rec["eno"].append(mc.energy_calculator.total_energy)
                a = mc.nf_big_move(cfg)
                rec["acc"].append(1 if a else 0)
                rec["E_after"].append(mc.energy_calculator.total_energy)
            key = "%s_s%d" % (tag, seed)
            names.append(key)
            out[key + "__pos0"] = p32
            out[key + "__seed"] = np.int64(seed)
            out[key + "__rounds"] = np.int64(rounds)
            out[key + "__local"] = np.int64(local)
            out[key + "__props"] = props.astype(np.float32)
            out[key + "__acc"] = np.array(rec["acc"], np.uint8)
            out[key + "__eno"] = np.array(rec["eno"], np.float64)
            out[key + "__E_after"] = np.array(rec["E_after"], np.float64)
            out[key + "__pos_before"] = np.array(rec["pos_before"], np.float64)
            out[key + "__idx"] = np.array(spy.ints, np.int64)
            out[key + "__u"] = np.array(spy.uniforms, np.float64)
            out[key + "__posF"] = np.asarray(mc.particles, dtype=np.float64)
            out[key + "__attempts"] = np.int64(mc.attempts_displacement)
            out[key + "__accepted"] = np.int64(mc.accepted_displacement)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "mc_global.npz"), **out)
    print("mc_global: %d traces" % len(names))


def gen_init(ref):
    """Fresh (unperturbed) reference flow for a fixed torch seed: pins parameter names,
    shapes and the construction order of the random initialisation."""
    NF = ref["normflows"]
    torch.manual_seed(123)
    with ref["quiet"]():
        base = NF.Energy.UniformParticle(3, 2, 5.0, device="cpu")
        layers = [NF.flows.CircularCoupledRationalQuadraticSpline(6, 2, 16, range(6), num_bins=8, tail_bound=5.0)
                  for _ in range(2)]
        model = NF.NormalizingFlow(base, layers)
    out = {"sd__" + k: v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, "flow_init_seed123.npz"), **out)
    print("flow_init: %d tensors" % len(out))


def gen_observables(ref):
    """Outputs of the unmodified hybrid_NF_MCMC/utils.py analysis functions on seeded configurations."""
    from oracle import _refimport
    hu = _refimport.load_hybrid_utils()
    rng = np.random.default_rng(77)
    out = {}
    # well statistics: N = 3 particles in an L = 10 box, a mix of all-in-A, all-in-B and scattered configurations
    half_box, r0, n, M = 5.0, 1.2, 3, 60
    L = 2 * half_box
    cfgs = np.empty((M, n, 2), dtype=np.float32)
    for m in range(M):
        kind = m % 3
        if kind == 2:
            cfgs[m] = rng.uniform(0, L, size=(n, 2))
        else:
            cx = L / 4 if kind == 0 else 3 * L / 4
            ang = rng.uniform(0, 2 * np.pi, n)
            rad = rng.uniform(0, 1.3, n)                      # a few land just outside 1.1 r0 = 1.32
            cfgs[m, :, 0] = cx + rad * np.cos(ang)
            cfgs[m, :, 1] = L / 2 + rad * np.sin(ang)
    cfgs[5, 0] = [L / 4 - L, L / 2 + L]                        # periodic images of the well centre
    cls = hu.classify_particles(cfgs, half_box, r0)
    code = np.zeros(cls.shape, dtype=np.uint8)
    code[cls == "A"] = 1
    code[cls == "B"] = 2
    avg_x, p_a, p_b, dF, runs = hu.calculate_well_statistics(cfgs, 7, half_box, r0)
    out.update(ws_cfgs=cfgs, ws_half_box=half_box, ws_r0=r0, ws_start=7, ws_class=code,
               ws_avg_x=np.array(avg_x, dtype=np.float64), ws_p_a=np.array(p_a), ws_p_b=np.array(p_b),
               ws_dF=np.array(dF), ws_runs=np.array(runs))
    # pair correlation: float32 centred samples, as model.sample(...).cpu().numpy() delivers them
    # (the reference's default dr = bound / 50 makes np.arange(0, bound + dr, dr) one bin too long for most
    # bounds and the function then fails on a shape mismatch; dr = bound / 49.5 keeps both aranges at 50)
    for tag, npart, nsamp, bound, dr in (("a", 8, 12, float(np.float32(np.sqrt(8 / 0.03))) / 2, None),
                                         ("b", 32, 6, float(np.float32(np.sqrt(32 / 0.03))) / 2, None),
                                         ("c", 3, 20, 5.0, "default")):
        if dr is None:
            dr = bound / 49.5
        samples = rng.uniform(-bound, bound, size=(nsamp, npart, 2)).astype(np.float32)
        samples[0, 1] = samples[0, 0]                          # coincident particles: zero distance is dropped
        with ref["quiet"]():
            if dr == "default":
                r_vals, g_r = hu.calculate_pair_correlation(samples, npart, bound)
                dr = bound / 50
            else:
                r_vals, g_r = hu.calculate_pair_correlation(samples, npart, bound, dr)
        out.update({"pc_%s_samples" % tag: samples, "pc_%s_bound" % tag: bound, "pc_%s_n" % tag: npart,
                    "pc_%s_dr" % tag: dr,
                    "pc_%s_r" % tag: np.asarray(r_vals), "pc_%s_g" % tag: np.asarray(g_r, dtype=np.float64)})
    np.savez_compressed(os.path.join(GOLD, "observables.npz"), **out)
    print("observables.npz:", len(out), "arrays")


def gen_judge(ref):
    """judge_normalizing_flow / bulk_judge_normalizing_flow (monte_carlo.py:305-370): energy-only Metropolis tests of
    proposals; the cached energy must be restored, attempts counted, one uniform consumed per finite uphill proposal."""
    SB = ref["simulation_box"].SimulationBox
    MC = ref["monte_carlo"].MonteCarlo
    lg = quiet_logger()
    out = {}
    n, rho, seed = 8, 0.3, 77
    L = er.box_length(n, rho)
    p32, _ = er.jittered_lattice(n, rho, seed)
    # float32 state from the start (what a reference chain holds after its first NF acceptance): the
    # device chain then follows the same trajectory bit for bit
    mc = MC(p32.copy(), SB(L), 1.0, n, initial_max_displacement=0.5, timing=False, checking=False,
            logger=lg, seed=seed, device=torch.device("cpu"), **POT)
    for _ in range(30):
        mc.particle_displacement()
    rng = np.random.default_rng(5)
    state = np.array(mc.particles, dtype=np.float32)
    props = []
    for i in range(24):                                   # small perturbations (mixed up / downhill), one overlap
        c = state + (rng.random((n, 2)).astype(np.float32) - 0.5) * np.float32(0.08 * (1 + i % 4))
        c = np.mod(c, np.float32(L)).astype(np.float32)
        props.append(c)
    props[7][1] = props[7][0] + np.float32(0.2)
    e_before = mc.energy_calculator.total_energy
    att0 = mc.attempts_displacement
    crit = [bool(mc.judge_normalizing_flow(c.copy())) for c in props[:12]]
    out.update(pos0=p32, L=np.float64(L), n=np.int64(n), seed=np.int64(seed), warm=np.int64(30),
               props=np.stack(props), crit=np.array(crit), e_cached=np.float64(mc.energy_calculator.total_energy),
               e_before=np.float64(e_before), att_delta=np.int64(mc.attempts_displacement - att0))
    ref_e = float(e_before) + 0.5
    acc, att = mc.bulk_judge_normalizing_flow([c.copy() for c in props[12:]], ref_e)
    out.update(bulk_ref_energy=np.float64(ref_e), bulk_acc=np.int64(acc), bulk_att=np.int64(att),
               next_uniform=np.float64(mc.rng.random()))
    np.savez_compressed(os.path.join(GOLD, "mc_judge.npz"), **out)
    print("mc_judge: %d single + %d bulk proposals, %d accepted" % (12, att, acc))


def gen_initialise(ref):
    """MCMC/initialise.py: initialise_low_left / right (N = 1..12) and initialise_fcc on a few shapes.  The module
    imports matplotlib for its plots; inert stand-ins are registered first (as for hybrid_NF_MCMC/utils.py)."""
    import importlib.util
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("fs_ref_initialise",
                                                  os.path.join(_refimport.REF_ROOT, "MCMC", "initialise.py"))
    ini = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ini)
    out = {}
    for n in range(1, 13):
        for rho in (0.03, 0.5):
            for side, fn in (("L", ini.initialise_low_left), ("R", ini.initialise_low_right)):
                p, box = fn(n, rho, 1.0)
                out["low%s_n%d_rho%g" % (side, n, rho)] = np.asarray(p, dtype=np.float64)
                out["low%s_n%d_rho%g_box" % (side, n, rho)] = np.array([box.box_size_x, box.box_size_y])
    p, box = ini.initialise_low_left(6, 0.5, 2.0)
    out["lowL_n6_rho0.5_ar2"] = np.asarray(p, dtype=np.float64)
    out["lowL_n6_rho0.5_ar2_box"] = np.array([box.box_size_x, box.box_size_y])
    for n, rho, ar in ((48, 0.5, 1.5), (32, 0.03, 1.0), (7, 0.3, 1.0), (256, 0.5, 1.0)):
        with ref["quiet"]():
            p, box = ini.initialise_fcc(n, rho, ar)
        out["fcc_n%d_rho%g_ar%g" % (n, rho, ar)] = np.asarray(p, dtype=np.float64)
        out["fcc_n%d_rho%g_ar%g_box" % (n, rho, ar)] = np.array([box.box_size_x, box.box_size_y])
    np.savez_compressed(os.path.join(GOLD, "initialise.npz"), **out)
    print("initialise.npz:", len(out), "arrays")


def gen_affine(ref):
    """RealNVP-style flows of the vendored normflows (flows/affine/coupling.py:99-268, flows/periodic.py:6-73,
    nets/mlp.py): a MaskedAffineFlow stack with PeriodicShift layers in between, AffineCouplingBlocks with the three scale
    maps, PeriodicWrap; layer by layer forward / inverse outputs and log-determinants, and the container's log_prob."""
    NF = ref["normflows"]
    torch.manual_seed(31)
    D, bound, B = 12, 4.5, 40
    g = torch.Generator().manual_seed(32)
    out = {"D": np.int64(D), "bound": np.float64(bound)}
    b = torch.tensor([1.0 if i % 2 == 0 else 0.0 for i in range(D)])
    flows = []
    for i in range(4):
        s_net = NF.nets.MLP([D, 2 * D, D], init_zeros=True)
        t_net = NF.nets.MLP([D, 2 * D, D], init_zeros=True)
        flows.append(NF.flows.MaskedAffineFlow(b if i % 2 == 0 else 1 - b, t_net, s_net))
        flows.append(NF.flows.PeriodicShift(list(range(0, D, 3)), bound=bound, shift=0.37 * (i + 1)))
    with ref["quiet"]():
        base = NF.Energy.UniformParticle(D // 2, 2, bound, device="cpu")
    model = NF.NormalizingFlow(base, flows)
    with torch.no_grad():
        for prm in model.parameters():
            prm.add_(0.04 * torch.randn(prm.shape, generator=g))
    model.eval()
    z = (torch.rand(B, D, generator=g) * 2 - 1) * bound
    x = (torch.rand(B, D, generator=g) * 2 - 1) * bound
    with torch.no_grad():
        zz, lds = z.clone(), []
        for i, f in enumerate(model.flows):
            zz, ld = f(zz)
            out["stack_fwd_%d" % i] = zz.numpy().copy()
            lds.append(ld.numpy().copy())
        out["stack_fwd_ld"] = np.stack(lds)
        xx, lds = x.clone(), []
        for i in range(len(model.flows) - 1, -1, -1):
            xx, ld = model.flows[i].inverse(xx)
            out["stack_inv_%d" % i] = xx.numpy().copy()
            lds.append(ld.numpy().copy())
        out["stack_inv_ld"] = np.stack(lds)
        out["stack_log_prob"] = model.log_prob(x.clone()).numpy()
        zi, ldi = model.inverse_and_log_det(x.clone())
        out["stack_inv_total_ld"] = ldi.numpy()
    out["stack_z"], out["stack_x"] = z.numpy(), x.numpy()
    for k, v in model.state_dict().items():
        out["stack_sd__" + k] = v.numpy().copy()          # (copy: the NaN case below edits a weight in place)
    # non-finite parameter -> NaN (coupling.py:199-202)
    with torch.no_grad():
        zbig = z.clone()
        zbig[0, 0] = 1e30
        model.flows[0].s.net[-1].weight.mul_(1e12)
        y, ld = model.flows[0](zbig)
        out["nan_in"], out["nan_out"], out["nan_ld"] = zbig.numpy(), y.numpy(), ld.numpy()
        out["nan_last_w"] = model.flows[0].s.net[-1].weight.numpy().copy()
    # AffineCouplingBlock with the three scale maps, both split modes
    for sm in ("exp", "sigmoid", "sigmoid_inv"):
        for mode in ("channel", "channel_inv"):
            pm = NF.nets.MLP([D // 2, 16, D], init_zeros=True)
            blk = NF.flows.AffineCouplingBlock(pm, scale=True, scale_map=sm, split_mode=mode)
            with torch.no_grad():
                for prm in blk.parameters():
                    prm.add_(0.08 * torch.randn(prm.shape, generator=g))
                yf, lf = blk(z.clone())
                yi, li = blk.inverse(x.clone())
            tag = "blk_%s_%s" % (sm, mode)
            out[tag + "_fwd"], out[tag + "_fwd_ld"], out[tag + "_inv"], out[tag + "_inv_ld"] = (
                yf.numpy(), lf.numpy(), yi.numpy(), li.numpy())
            for k, v in blk.state_dict().items():
                out[tag + "_sd__" + k] = v.numpy()
    pm = NF.nets.MLP([D // 2, 16, D // 2], init_zeros=True)
    blk = NF.flows.AffineCouplingBlock(pm, scale=False)
    with torch.no_grad():
        for prm in blk.parameters():
            prm.add_(0.2 * torch.randn(prm.shape, generator=g))
        yf, lf = blk(z.clone())
    out["blk_noscale_fwd"], out["blk_noscale_fwd_ld"] = yf.numpy(), lf.numpy()
    for k, v in blk.state_dict().items():
        out["blk_noscale_sd__" + k] = v.numpy()
    wrap = NF.flows.PeriodicWrap(list(range(1, D, 2)), bound=bound)
    far = z * 3.0
    with torch.no_grad():
        out["wrap_in"], out["wrap_inv"] = far.numpy(), wrap.inverse(far.clone())[0].numpy()
    np.savez_compressed(os.path.join(GOLD, "affine.npz"), **out)
    print("affine.npz:", len(out), "arrays")


def gen_target(ref):
    """Training target of Algorithm 2: NF.Energy.DoubleWellLJ._energy (NF/normflows/Energy/SimpleLJ.py:42-128) and its
    gradient by the reference's own autograd, on float32 centred configurations incl. soft-core pairs (r <= 0.82),
    particles near the origin particle, in the wells, and outside the box (wrapped by the reference)."""
    NF = ref["normflows"]
    out = {}
    names = []
    for tag, n, rho, T in (("n3", 3, 0.03, 1.0), ("n8", 8, 0.1, 1.0), ("n64", 64, 0.03, 0.7)):
        bound = er.box_length(n, rho) / 2
        tgt = NF.Energy.DoubleWellLJ(2 * n, n, T, bound, V0_list=[-10.0, -10.5], r0=1.2, k=15)
        g = torch.Generator().manual_seed(17 + n)
        B = 24
        x = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
        x[0, :4] = torch.tensor([0.3, 0.1, 0.75, 0.2])            # soft-core pair and a particle next to the origin one
        x[1, :2] = torch.tensor([-bound / 2 + 1.1, 0.2])           # on the wall of the left well
        x[2, :2] = torch.tensor([bound / 2, 0.05])                 # inside the right well
        x[3, 0] = bound * 1.3                                      # outside the box: wrapped (:19-20)
        x[4, :4] = torch.tensor([bound - 0.2, 1.0, -bound + 0.2, 1.0])   # neighbours across the boundary: NO minimum image
        xr = x.clone().requires_grad_(True)
        with _refimport.cuda_zeros_on_cpu():
            E = tgt._energy(xr)
            (grad,) = torch.autograd.grad(E.sum(), xr)
            lj = NF.Energy.SimpleLJ._energy(tgt, x.clone())
            dw = tgt.double_well_potential(x.clone().view(B, n, 2))
        out.update({tag + "__x": x.numpy(), tag + "__E": E.detach().numpy(), tag + "__grad": grad.numpy(),
                    tag + "__lj": lj.numpy(), tag + "__dw": dw.numpy(), tag + "__bound": np.float64(bound),
                    tag + "__T": np.float64(T), tag + "__n": np.int64(n)})
        names.append(tag)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "target_energy.npz"), **out)
    print("target_energy: %d cases" % len(names))


def main():
    """No arguments: every fixture.  `flow:<tag>[,<tag>]`: only those flow fixtures."""
    ref = _refimport.load()
    os.makedirs(GOLD, exist_ok=True)
    sel = [a for a in sys.argv[1:] if a.startswith("flow:")]
    if sel:
        gen_flow(ref, only=set(sel[0][5:].split(",")))
        return
    if "target" in sys.argv[1:]:
        gen_target(ref)
        return
    if "judge" in sys.argv[1:]:
        gen_judge(ref)
        return
    if "initialise" in sys.argv[1:]:
        gen_initialise(ref)
        return
    if "affine" in sys.argv[1:]:
        gen_affine(ref)
        return
    gen_energy(ref)
    gen_mc(ref)
    gen_flow(ref)
    gen_global(ref)
    gen_init(ref)
    gen_observables(ref)
    gen_target(ref)
    gen_judge(ref)
    gen_initialise(ref)
    gen_affine(ref)


if __name__ == "__main__":
    main()
