#!/usr/bin/env python
"""Benchmark of the NF-MCMC sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload alg1_n256|alg1_n32|alg2_n64]
                    [--precision auto|tf32|fp32] [--impl ours|reference] [--no-secondary]

One "step" = one hybrid Algorithm-1 round on every chain of this rank
(hybrid_NF_MCMC/main_algorithm_1.py:381-395): BIG_MOVE_INTERVAL local
displacement moves + ONE NF-proposed global move (flow sample, total energy of
the proposal, log q of old and new state, Metropolis accept).  The metric is MH
chain-steps/s (every local move and every global move is one chain-step, like
the reference's attempts counter); NF proposals + energy evals/s is reported
beside it.

Default workload = the north_star target, BASELINE configs[2]: Algorithm-1 hybrid
at N = 256 particles, 8192 chains per GPU (65 536 over 8 GPUs), flow K=15 / H=256 /
32 blocks / 32 bins, for every --gpus N (weak scaling: chains shard over ranks
with no data-path collective, flow weights are broadcast once from rank 0).
At N = 1 the same JSON line carries `secondary` entries for configs[1]
(alg1_n32), configs[3] (alg2_n64: the WHOLE Algorithm-2 cycle - local moves,
training step with gradient all-reduce, device-side re-pack, global move) and configs[4] (energy sweep
N = 64..4096 against a FP32-FMA peak measured in the same run), each with its
own roofline, plus the accept kernel's HBM figure.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: particles, chains per GPU, local steps per round, flow (K, blocks, H, bins), sigma, rho
    "alg1_n256": dict(n=256, chains=8192, local=1000, K=15, blocks=32, H=256, nb=32, sigma=0.02, rho=0.03,
                      desc="BASELINE configs[2]: Alg 1 hybrid, N=256, 8192 chains/GPU (65536 over 8 GPUs)"),
    "alg1_n32": dict(n=32, chains=4096, local=1000, K=15, blocks=32, H=256, nb=32, sigma=0.02, rho=0.03,
                     desc="BASELINE configs[1]: Alg 1 hybrid, 4096 chains/GPU, N=32"),
    # Algorithm 2 (main_algorithm_2.py:393-577): every cycle = 100 local moves per chain sampled every 10 steps,
    # one epoch of training on UPDATE_NUM_SAMPLES = 1000 of the new configurations (4 minibatches of 256, new Adam,
    # lr 5.435e-4, wd 9.586e-5), eval -> device-side re-pack, one NF global move per chain.  Under N GPUs every rank
    # trains on its own 1000 samples and the flat gradient bucket is all-reduced (NCCL, SUM / N) once per optimizer step.
    "alg2_n64": dict(n=64, chains=4096, local=100, K=23, blocks=2, H=128, nb=15, sigma=0.05, rho=0.03,
                     train=dict(samples=1000, batch=256, lr=0.000543510751759681, wd=9.5857178422352e-05, every=10),
                     desc="BASELINE configs[3]: Alg 2 whole cycle, N=64: 100 local steps + training step (4 x 256, "
                          "gradient all-reduce) + re-pack + 1 global move"),
}
DEFAULT_WORKLOAD = "alg1_n256"
POT = dict(num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15)
PHILOX_SEED = 20261018


def workload_config(name, w):
    """The `config` object both arms print (same keys, same values)."""
    return {"workload": name, "desc": w["desc"], "chains_per_gpu": w["chains"], "particles": w["n"],
            "local_steps_per_round": w["local"], "rho": w["rho"],
            "flow": {"K": w["K"], "blocks": w["blocks"], "H": w["H"], "bins": w["nb"], "sigma": w["sigma"]},
            "potential": {"wells": POT["num_wells"], "V0": POT["V0_list"], "r0": POT["r0"], "k": POT["k"], "T": 1.0},
            "l2": "inputs larger than L2: the flow weights (%.0f MB FP32) are streamed in every pass"
                  % (flow_params(w) * 4 / 1e6),
            "train": w.get("train")}


def flow_params(w):
    n, H, P = w["n"], w["H"], 3 * w["nb"] + 1
    per = (2 * n * H + H) + w["blocks"] * (2 * (H * H + H) + 4 * H) + (n * P * H + n * P) + n * (3 * w["nb"] + 1) + 2 * n
    return w["K"] * per


def build_flow(NF, w, bound, device, seed=0):
    """Random-init flow of the named architecture, perturbed so it is not the identity
    (SURVEY.md 8d): params += N(0, sigma^2), BN running stats randomised; eval mode."""
    torch.manual_seed(seed)
    base = NF.Energy.UniformParticle(w["n"], 2, bound, device=device)
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * w["n"], w["blocks"], w["H"], range(2 * w["n"]),
                                                              num_bins=w["nb"], tail_bound=bound)
              for _ in range(w["K"])]
    model = NF.NormalizingFlow(base, layers)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(w["sigma"] * torch.randn(p.shape, generator=g))
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    return model


def flops_per_sample_layer(w):
    """2 (2N H + 2 n_blocks H^2 + H N (3 nb + 1))."""
    return 2.0 * (2 * w["n"] * w["H"] + 2 * w["blocks"] * w["H"] ** 2 + w["H"] * w["n"] * (3 * w["nb"] + 1))


def sweep_flops_per_step(n):
    """old and new particle energy, E and W both tracked."""
    return 27.0 * 2 * (n - 1) + 2 * 40.0


def energy_flops(B, n):
    return 27.0 * B * n * (n - 1) / 2 + 40.0 * B * n


class quiet_host:
    """Timed regions run with the cyclic garbage collector paused (collected right before): a generation-2 collection
    of this process's heap takes 60-130 ms (measured: one round of 3.4 ms took 86 ms, always a single round of a run),
    during which no kernel is submitted and the GPU runs dry.  Reference counting still frees everything the rounds
    allocate; nothing in the timed loop creates cycles."""

    def __enter__(self):
        import gc
        gc.collect()
        self._was = gc.isenabled()
        gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()
        return False


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md).  Sampled in-process through NVML
    (nvidia_ml_py, the library nvidia-smi itself uses) from a background thread every 25 ms: an external
    `nvidia-smi -lms` poller re-initialises NVML over every GPU of the box and was measured to stall kernel submission
    of all ranks (2 GPUs: 3.87 -> 4.5-7.6 ms per round).  Falls back to one nvidia-smi query if NVML cannot be loaded."""

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = False
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            import threading
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(local_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if local_index < len(ids) and ids[local_index].strip().isdigit():
                return int(ids[local_index])
        return local_index

    def _sample(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop:
            try:
                self._sample()
            except Exception:
                return
            time.sleep(0.025)

    def wait_ready(self, timeout=2.0):
        t0 = time.time()
        while self._thread is not None and not self.sm and time.time() - t0 < timeout:
            time.sleep(0.005)

    def mark(self):
        """Start of the timed region: samples taken before this index are warm-up."""
        self._mark = len(self.sm)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self._thread is not None:
            self._stop = True
            self._thread.join(timeout=2)
        elif self._nvml is None:
            try:                                       # one query after the timed region (no NVML bindings available)
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                c = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                   timeout=20).stdout.strip().split(",")
                self.sm.append(float(c[0]))
                self.mx.append(float(c[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
            out["sm_max_mhz"] = float(max(self.mx))
            out["samples"] = len(self.sm)
        out["reasons"] = sorted(self.reasons)
        return out


# ---------------------------------------------------------------------------
# CPU side (the reference arm and the cpu_baseline leg): oracle port on the host cores
# ---------------------------------------------------------------------------
def _cpu_local_worker(args):
    n, rho, seed, steps = args
    from oracle import energy_ref as er
    from oracle import mc_ref as mr
    pot = er.Potential(POT["num_wells"], POT["V0_list"], POT["r0"], POT["k"])
    pos, L = er.jittered_lattice(n, rho, seed)
    ch = mr.ChainRef(pos.copy(), L, 1.0, pot, 0.65, rng=np.random.default_rng(seed))
    for _ in range(20):
        ch.local_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ch.local_step()
    return time.perf_counter() - t0


def cpu_baseline(w, budget_s=20.0, state_dict=None, bound=None):
    """Times the oracle on the host cores: `cores` processes each advance one chain by a bounded
    number of local steps; one batch of global moves (flow sample + 2 log-densities + total
    energy) runs on all torch threads.  Extrapolated to the hybrid round of the workload."""
    import multiprocessing as mp
    from oracle import energy_ref as er
    from oracle import flow_ref as fr
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    # local moves
    t_probe = _cpu_local_worker((w["n"], w["rho"], 1, 50)) / 50
    steps = int(max(50, min(5000, 0.5 * budget_s / max(t_probe, 1e-6))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_local_worker, [(w["n"], w["rho"], 100 + i, steps) for i in range(cores)])
    t_local = float(np.mean(times)) / steps                    # seconds per chain-step per core
    # global moves
    bsz = 16
    spec = fr.FlowSpec(state_dict, bound)
    pot = er.Potential(POT["num_wells"], POT["V0_list"], POT["r0"], POT["k"])
    L = 2 * bound
    g = torch.Generator().manual_seed(0)
    z = (torch.rand(bsz, 2 * w["n"], generator=g) * 2 - 1) * bound
    old = (torch.rand(bsz, 2 * w["n"], generator=g) * 2 - 1) * bound
    t0 = time.perf_counter()
    with torch.no_grad():
        x, _ = fr.forward_and_log_det(state_dict, spec, z)
        fr.log_prob(state_dict, spec, torch.cat([old, x]))
    cfg = (x.numpy() + np.float32(bound)).reshape(bsz, w["n"], 2)
    for b in range(bsz):
        er.total_energy_virial(cfg[b], L, L, pot)
    t_global = (time.perf_counter() - t0) / bsz               # seconds per proposal (all cores)
    round_s = w["local"] * t_local / cores + t_global           # per chain, all cores busy
    value = (w["local"] + 1) / round_s
    return {"value": value, "unit": "chain-steps/s", "cores": cores, "kind": "port",
            "sample": "%d procs x %d local steps (N=%d) + %d global moves (flow K=%d H=%d blocks=%d on %d torch "
                      "threads), extrapolated to %d local + 1 global per chain"
                      % (cores, steps, w["n"], bsz, w["K"], w["H"], w["blocks"], cores, w["local"]),
            "local_steps_per_s_per_core": 1.0 / t_local, "global_moves_per_s": 1.0 / t_global}


def parity_check(path):
    """The checker half of `parity_checked` (runs in the CPU subprocess, oracle only): re-evaluates a small sample of
    what the GPU arm computed right before its timed region - a traced local sweep (lock-step, every decision), the
    total energy of proposals, the flow in both directions and the global accept rule - and returns the errors."""
    from oracle import energy_ref as er
    from oracle import flow_ref as fr
    from oracle import mc_ref as mr
    from oracle import philox_ref as pr
    d = np.load(path + ".npz")
    sd = torch.load(path + ".pt", map_location="cpu")
    n, L, md = int(d["n"]), float(d["L"]), float(d["md"])
    bound = L / 2
    pot = er.Potential(POT["num_wells"], POT["V0_list"], POT["r0"], POT["k"])
    out = {"decisions": 0, "flips_in_band": 0, "outside_band": 0, "sweep_energy_err": 0.0}
    steps = d["acc"].shape[1]
    for c in range(d["pos0"].shape[0]):
        p_all, u_all = pr.step_draws(int(d["seed"]), int(d["chain_id0"]) + c, 0, steps, n)
        r = mr.lockstep_check(d["pos0"][c], L, md, pot, p_all, u_all, d["acc"][c], d["idx"][c], d["e"][c])
        out["decisions"] += steps
        out["flips_in_band"] += r["flips_in_band"]
        out["outside_band"] += r["outside_band"] + int(not np.array_equal(r["final"], d["posF"][c]))
        out["sweep_energy_err"] = max(out["sweep_energy_err"], r["max_energy_err"])
    spec = fr.FlowSpec(sd, bound)
    with torch.no_grad():
        xo, ldo = fr.forward_and_log_det(sd, spec, torch.from_numpy(d["z"]).double(), dtype=torch.float64)
        lq = fr.log_prob(sd, spec, torch.from_numpy(d["lq_in"]).double(), dtype=torch.float64).numpy()
    out["sample_err_of_bound"] = float((torch.from_numpy(d["x"]).double() - xo).abs().max() / bound)
    out["logq_rel_err"] = float(np.max(np.abs(d["lq"] - lq) / np.abs(lq)))
    cfg = d["cfg"]
    e_err = 0.0
    for b in range(cfg.shape[0]):
        Er, _ = er.total_energy_virial(cfg[b].astype(np.float64), L, L, pot)
        if np.isfinite(Er):
            e_err = max(e_err, abs(float(d["E_new"][b]) - Er) / max(1.0, abs(Er)))
        else:
            e_err = max(e_err, 0.0 if np.isinf(d["E_new"][b]) else 1.0)
    out["energy_rel_err"] = e_err
    # global accept rule (monte_carlo.py:264-303) on the device's own inputs
    B = cfg.shape[0]
    lq_old, lq_new = d["lq"][:B].astype(np.float64), d["lq"][B:].astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        ratio = np.exp(-(d["E_new"].astype(np.float64) - d["E_old"]) - ((-lq_new) - (-lq_old)))
    ref_mask = (ratio >= 1.0) | (d["u"] < ratio)
    out["accept_mismatch"] = int(np.sum(ref_mask != d["mask"].astype(bool)))
    out["ok"] = bool(out["outside_band"] == 0 and out["sweep_energy_err"] < 1e-5 and out["logq_rel_err"] < 1e-4
                     and out["energy_rel_err"] < 1e-5 and out["accept_mismatch"] == 0
                     and out["sample_err_of_bound"] < 2e-4)
    return out


def run_reference(args, name, w):
    """--impl reference: the reference's CPU path.  The reference is pure Python and does not
    travel to the GPU box (nothing to compile into oracle/_ref), so this arm times the oracle
    port of it on all host cores, on the same config / metric / unit.  Nothing of the product
    package is imported here: the flow weights come from oracle/flow_init.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import flow_init
    bound = float(np.float32(np.sqrt(w["n"] / w["rho"]))) / 2
    sd = flow_init.synthetic_state_dict(w["n"], w["K"], w["blocks"], w["H"], w["nb"], w["sigma"], seed=0)
    vals = []
    base = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(w, budget_s=args.budget, state_dict=sd, bound=bound)
        if i >= args.warmup:
            vals.append(base["value"])
        if time.perf_counter() - t_all > 200:
            break
    v = float(np.mean(vals)) if vals else base["value"]
    base["value"] = v
    if args.parity_file:
        try:
            base["parity"] = parity_check(args.parity_file)
        except Exception as ex:
            base["parity"] = {"ok": False, "error": repr(ex)}
    line = {"impl": "reference", "metric": "mh_chain_steps_per_s", "value": v, "unit": "chain-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (w["local"] + 1) * w["chains"] * args.gpus / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, w),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
def load_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        return json.load(open(pk)), "measured (MEASURED_PEAKS.json)"
    # fallback of /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6550.0, "bf16_tflops": 1630.0, "bf16_tflops_sustained": 1375.0}, "fallback (B200_PROFILING.md)"


def fp32_peak(device_index):
    """FP32 FMA peak measured in this run (scripts/fp32_peak.cu, built by __graft_entry__.build())."""
    so = os.path.join(ROOT, "scripts", "libfp32_peak.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)
    lib.fp32_peak.restype = ctypes.c_int
    lib.fp32_peak.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    out = {}
    for variant, tag in ((0, "ffma"), (1, "ffma2")):
        tf, ms = ctypes.c_double(), ctypes.c_double()
        if lib.fp32_peak(device_index, variant, ctypes.byref(tf), ctypes.byref(ms)) != 0:
            return None
        out[tag] = tf.value
    out["peak"] = max(out["ffma"], out["ffma2"])
    return out


def time_graphed_round(h, steps):
    """The same round issued as ONE CUDA-graph launch (flowstate_b200/drivers/rounds.HybridRound): a reported variant, the
    headline stays the host-driven loop.  Returns None when the capture is refused."""
    from flowstate_b200.drivers.rounds import HybridRound
    try:
        hr = HybridRound(h.eng, h.model, h.w["local"], use_graph=True)
        with quiet_host():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(6):
                hr.step()
            b.record()
            torch.cuda.synchronize()
            per = max(a.elapsed_time(b) / 6, 1e-3)
            for _ in range(max(0, min(200, int(200.0 / per) - 6))):
                hr.step()
            torch.cuda.synchronize()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            marks[0].record()
            for i in range(steps):
                hr.step()
                marks[i + 1].record()
            torch.cuda.synchronize()
        ms = marks[0].elapsed_time(marks[-1])
        per_round = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        return {"value": h.B * (h.w["local"] + 1) * steps / (ms * 1e-3), "unit": "chain-steps/s",
                "ms_per_step": ms / steps, "steps": steps, "graph_launches_per_step": 1,
                "kernels_per_graph": hr.launches_per_round,
                "step_ms_min_median_max": [min(per_round), sorted(per_round)[len(per_round) // 2], max(per_round)],
                "note": "whole round (forked sampling pass, sweep, log-densities, fused global move) replayed from a CUDA "
                        "graph; bit-equal to the eager round: tests/test_gpu_drivers.py::"
                        "test_graphed_hybrid_round_equals_eager_round"}
    except Exception as ex:
        return {"error": repr(ex)[:200]}


class Harness:
    """One workload on this rank: engine + flow + the hybrid round, its timings and rooflines."""

    def __init__(self, name, w, dev, rank, world, precision):
        import flowstate_b200.MCMC as MC
        import flowstate_b200.normflows as NF
        from flowstate_b200 import _lib, parallel
        self.MC, self.NF, self._lib, self.parallel = MC, NF, _lib, parallel
        self.name, self.w, self.dev, self.rank, self.world = name, w, dev, rank, world
        n, B = w["n"], w["chains"]
        self.n, self.B = n, B
        self.L = float(np.float32(np.sqrt(n / w["rho"])))
        self.bound = self.L / 2
        # flow: built on every rank from the same seed on CPU, then broadcast from rank 0 over NCCL
        self.model = build_flow(NF, w, self.bound, dev).to(dev).eval()
        self.bcast_bytes = parallel.broadcast_flow(self.model, src=0) if world > 1 else 0
        prec = precision
        if prec == "auto":
            prec = "tf32"
            try:
                self.model.precision = "tf32"
                self.model.log_prob(torch.zeros(2, 2 * n, device=dev))
            except Exception:
                prec = "fp32"
        self.prec = prec
        self.model.precision = prec
        self.fused = prec == "tf32" and w["H"] in (128, 256) and w["nb"] <= 32 and not os.environ.get("FS_NO_FUSE")
        # chains: this rank's contiguous block of global chain ids
        self.start, _ = parallel.shard_range(B * world, rank, world)
        pos, _ = MC.jittered_lattice(n, w["rho"], seed=1000 + self.start, batch=B)
        self.pos0 = pos
        self.eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(self.L), 1.0, n, initial_max_displacement=0.65,
                                        rng="philox", philox_seed=PHILOX_SEED, chain_id0=self.start, device=dev, **POT)
        self.eng.set_nf_model(self.model)
        self.half32 = np.float32(self.bound)
        torch.manual_seed(1234 + rank)
        # Proposals do not depend on the chains (Alg 1 even pre-generates its whole pool,
        # main_algorithm_1.py:340-343), so the sampling pass of round r+1 runs on a side stream in the shadow
        # of round r's sweep + log-density pass; every round still does one sample pass, two log-densities,
        # one proposal energy and `local` local moves per chain.
        self.side = torch.cuda.Stream(device=dev)
        self.pending = {}
        self.logq_from_sample = False      # True: the reported VARIANT (details.variant_logq_from_sampling_pass)
        self.trainer = None
        if w.get("train"):
            from flowstate_b200.drivers.training import FlowTrainer
            t = w["train"]
            self.trainer = FlowTrainer(self.model, t["lr"], t["wd"], 1.0, t["batch"], use_graph=True)
            self.model.layer_parallel = "prefer"   # the cycle's sample / log_prob passes run alone (no side stream)

    # -- the round ----------------------------------------------------------
    def launch_proposals(self, z_host=None):
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.side):
            if z_host is None:
                z = self.model.q0(self.B)
            else:
                z = torch.empty(self.B, 2 * self.n, dtype=torch.float32, device=self.dev)
                z.copy_(z_host, non_blocking=True)
            lq_new = None
            if self.logq_from_sample:                  # variant: log q(proposal) from the sampling pass itself
                x, ld = self.model.forward_and_log_det(z)
                lq_new = self.model.q0.log_prob(z) - ld
                lq_new.record_stream(main)
            else:
                x = self.model.forward(z)
            cfg = (x.reshape(self.B, self.n, 2) + self.half32).contiguous()
            ev = torch.cuda.Event()
            ev.record(self.side)
        cfg.record_stream(main)
        self.pending["cfg"], self.pending["ev"], self.pending["lq_new"] = cfg, ev, lq_new

    def train_cycle(self, z_host=None, timers=None):
        """One Algorithm-2 cycle (main_algorithm_2.py:393-548)."""
        w, eng, model, t = self.w, self.eng, self.model, self.w["train"]

        def mark(name):
            if timers is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                timers.append((name, ev))
        mark("start")
        snaps = []
        for s0 in range(0, w["local"], t["every"]):            # SAMPLING_FREQUENCY = 10 (main_algorithm_2.py:59, 403-412)
            eng.particle_displacement(min(t["every"], w["local"] - s0))
            snaps.append(eng.centred(eng.pos))
        pool = torch.cat(snaps)
        data = pool[torch.randint(0, pool.shape[0], (t["samples"],), device=self.dev)]
        mark("local_sweep+collect")
        model.train()
        self.trainer.fresh_optimizer()                         # new Adam every cycle (main_algorithm_2.py:440)
        perm = torch.randperm(data.shape[0], device=self.dev)
        for bi in range(0, data.shape[0], t["batch"]):
            self.trainer.step(data[perm[bi:bi + t["batch"]]], sync=False)   # loss and skip decision stay on the device
        model.eval()
        mark("train")
        if self.world > 1:
            # parameters are identical by construction (all-reduced gradients): only rank 0's BatchNorm statistics move
            self.parallel.broadcast_flow(model, src=0, buffers_only=True)
        mark("broadcast")
        model._cuda_pack()                                     # device-side re-pack (fs_flow_update)
        mark("repack")
        if z_host is None:
            z = model.q0(self.B)
        else:
            z = torch.empty(self.B, 2 * self.n, dtype=torch.float32, device=self.dev)
            z.copy_(z_host, non_blocking=True)
        cfg = (model.forward(z).reshape(self.B, self.n, 2) + self.half32).contiguous()
        mark("flow_sample")
        mask = eng.nf_big_move(cfg)
        mark("global_move")
        return mask

    def one_round(self, z_host=None):
        if self.trainer is not None:
            return self.train_cycle(z_host)
        if "cfg" not in self.pending:
            self.launch_proposals(z_host)
        cfg, ev, lq_new = self.pending.pop("cfg"), self.pending.pop("ev"), self.pending.pop("lq_new")
        self.launch_proposals(z_host)                  # next round's proposals, off the critical path
        self.eng.particle_displacement(self.w["local"])
        torch.cuda.current_stream(self.dev).wait_event(ev)
        return self.eng.nf_big_move(cfg, logq_new=lq_new)

    # -- parity sample (checked by the oracle in the cpu_baseline subprocess) ---
    def dump_parity_sample(self, path, chains=6, steps=48, rows=6):
        MC = self.MC
        small = MC.BatchedMonteCarlo(self.pos0[:chains], MC.SimulationBox(self.L), 1.0, self.n,
                                     initial_max_displacement=0.65, rng="philox", philox_seed=PHILOX_SEED,
                                     chain_id0=self.start, device=self.dev, **POT)
        small.set_nf_model(self.model)
        tr = small.particle_displacement(steps, trace=True)
        posF = small.pos.clone()
        g = torch.Generator().manual_seed(99)
        z = ((torch.rand(chains, 2 * self.n, generator=g) * 2 - 1) * self.bound).to(self.dev)
        x, _ = self.model.forward_and_log_det(z)
        cfg = (x.reshape(chains, self.n, 2) + self.half32).contiguous()
        lq_in = torch.cat([small.centred(small.pos), small.centred(cfg)])
        lq = self.model.log_prob(lq_in)
        E_new, _, _ = small.total_energy_virial(cfg)
        E_old = small.E.clone()
        u = torch.rand(chains, generator=g, dtype=torch.float64).to(self.dev)
        mask = small.nf_big_move(cfg, u=u, logq=(lq[:chains].contiguous(), lq[chains:].contiguous()))
        c = lambda t: t.detach().cpu().numpy()
        np.savez(path + ".npz", n=self.n, L=self.L, md=0.65, seed=PHILOX_SEED, chain_id0=self.start,
                 pos0=self.pos0[:chains], acc=c(tr["accept"]), idx=c(tr["idx"]), e=c(tr["e"]), posF=c(posF), z=c(z),
                 x=c(x), cfg=c(cfg), lq_in=c(lq_in), lq=c(lq), E_new=c(E_new), E_old=c(E_old), u=c(u), mask=c(mask))
        torch.save({k: v.cpu() for k, v in self.model.state_dict().items()}, path + ".pt")

    # -- device-resident timing ---------------------------------------------
    def time_device(self, steps, warmup, clocks=None):
        import torch.distributed as dist
        dev, world = self.dev, self.world
        if world > 1:
            dist.barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # The garbage collection of quiet_host (60-130 ms of host time with the GPU idle) happens BEFORE the warm-up
        # rounds, not between them and the timed region.  (What remains at the start of a timed region is the refill of
        # the host-driven pipeline after the bracketing synchronize: until the host is a few rounds ahead again, the
        # side stream's proposals and the main stream's passes do not overlap - traced on the N = 32 workload: four
        # rounds of 3.1-4.6 ms before the steady 2.33 ms; part of the measurement, amortised over the timed rounds.)
        with quiet_host():
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            for _ in range(warmup):
                self.one_round()
            # ... and at least ~0.2 s of warm-up load for workloads with short rounds.  Same count on every rank.
            w1.record()
            torch.cuda.synchronize()
            per = max(w0.elapsed_time(w1) / max(warmup, 1), 1e-3)
            extra = torch.tensor([max(0, min(200, int(200.0 / per) - warmup))], device=dev)
            if world > 1:
                dist.all_reduce(extra, op=dist.ReduceOp.MAX)
            for _ in range(int(extra.item())):
                self.one_round()
            self.warmup_rounds = warmup + int(extra.item())
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = self._lib.lib().fs_launch_count()
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                self.one_round()
                marks[i].record()
            e1.record()
            torch.cuda.synchronize()
        launches = int(self._lib.lib().fs_launch_count() - l0)
        # per-round times (main stream), for the record: the headline stays e0 -> e1 over exactly `steps` rounds
        self.step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(steps)]
        if os.environ.get("FS_BENCH_TRACE"):
            print("[bench] %s per-round ms: %s" % (self.name, ["%.2f" % t for t in self.step_ms]), file=sys.stderr)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches

    # -- end to end: host buffers in, host results out, copies inside the timed region ----
    def time_e2e(self, steps):
        """Every step uploads that step's chain positions and base noise from pinned host memory, runs the round and
        reads positions, energies and accept mask back.  Two host-resident chain batches alternate (batch k % 2 in
        step k, its output becoming its next input), so the host waits for the results of step k-1 while step k is
        already queued: the copies and the launch work of consecutive steps overlap as in a real pipeline."""
        import torch.distributed as dist
        B, n, dev, eng = self.B, self.n, self.dev, self.eng
        h_z = torch.empty(B, 2 * n, dtype=torch.float32).uniform_(-self.bound, self.bound).pin_memory()
        # Copies ride on their own streams (the copy engines run beside the SMs): the upload of step k + 1's chain batch
        # is issued while step k computes, the download of step k's results while step k + 1 computes; device-side
        # staging buffers decouple them from the engine's state.  Every step still moves its own inputs and results.
        main = torch.cuda.current_stream(dev)
        up, dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        hb = []
        for k in range(2):
            hp = torch.empty(B, n, 2, dtype=torch.float32).pin_memory()
            hp.copy_(eng.pos.cpu())
            hb.append({"pos": hp, "out_pos": torch.empty(B, n, 2, dtype=torch.float32).pin_memory(),
                       "out_E": torch.empty(B, dtype=torch.float64).pin_memory(),
                       "out_mask": torch.empty(B, dtype=torch.uint8).pin_memory(),
                       "d_in": torch.empty(B, n, 2, dtype=torch.float32, device=dev),
                       "d_pos": torch.empty(B, n, 2, dtype=torch.float32, device=dev),
                       "d_E": torch.empty(B, dtype=torch.float64, device=dev),
                       "d_mask": torch.empty(B, dtype=torch.uint8, device=dev),
                       "ev_up": None, "ev_used": None, "ev_dn": None})

        def upload(k):
            buf = hb[k % 2]
            if buf["ev_dn"] is not None:                   # results of this batch's previous step have landed:
                buf["ev_dn"].synchronize()                 # its output buffer becomes its next input
                buf["pos"], buf["out_pos"] = buf["out_pos"], buf["pos"]
            with torch.cuda.stream(up):
                if buf["ev_used"] is not None:
                    up.wait_event(buf["ev_used"])          # the engine has taken the previous content of the staging buffer
                buf["d_in"].copy_(buf["pos"], non_blocking=True)
                buf["ev_up"] = torch.cuda.Event()
                buf["ev_up"].record(up)

        def e2e_round(k):
            buf = hb[k % 2]
            main.wait_event(buf["ev_up"])
            eng.pos.copy_(buf["d_in"])
            buf["ev_used"] = torch.cuda.Event()
            buf["ev_used"].record(main)
            eng.refresh_energy()
            mask = self.one_round(h_z)                     # base noise comes from the host buffer (side stream)
            if buf["ev_dn"] is not None:
                main.wait_event(buf["ev_dn"])              # the previous download out of the staging buffers is done
            buf["d_pos"].copy_(eng.pos)
            buf["d_E"].copy_(eng.E)
            buf["d_mask"].copy_(mask)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(dn):
                dn.wait_event(done)
                buf["out_pos"].copy_(buf["d_pos"], non_blocking=True)
                buf["out_E"].copy_(buf["d_E"], non_blocking=True)
                buf["out_mask"].copy_(buf["d_mask"], non_blocking=True)
                buf["ev_dn"] = torch.cuda.Event()
                buf["ev_dn"].record(dn)
            upload(k + 1)                                  # after this step is queued: the host wait inside it (step
                                                           # k - 1's results) overlaps this step's execution

        upload(0)
        with quiet_host():
            for k in range(4):
                e2e_round(k)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                e2e_round(k)
            torch.cuda.synchronize()
            e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        h2d = B * n * 2 * 4 + B * 2 * n * 4
        d2h = B * n * 2 * 4 + B * 8 + B
        return float(e2e_s.item()), h2d, d2h

    # -- dominant kernel, timed alone ------------------------------------------
    def conditioner_roofline(self, peaks, peak_src):
        """One coupling layer (conditioner GEMM chain with the fused spline epilogue, fs_flow_coupling), one launch per
        layer and pass, timed alone with CUDA events on the launching stream on the row count of the log-density pass.
        Timed in isolation -> the burst BF16 figure is the denominator (FP16 operands run at the BF16 rate, the TF32
        GEMM0 at half of it: flop-weighted harmonic blend)."""
        w, eng, model, dev = self.w, self.eng, self.model, self.dev
        xin = eng.centred(torch.cat([eng.pos, eng.pos]))
        model.log_prob(xin)
        pack = model._cuda_pack()
        s = np.pi / self.bound
        feats = torch.cat([torch.cos(xin[:, 0::2] * s), torch.sin(xin[:, 0::2] * s)], dim=1).contiguous()
        xo, ldo = torch.zeros_like(xin), torch.zeros(xin.shape[0], device=dev)
        feats_t = pack.tile_features(feats) if self.fused else None   # the layout the full passes hand to the kernel

        def dominant(li):
            if self.fused:
                pack.coupling(li, "density", feats_t, xin, xo, ldo, tiled=True)
            else:
                pack.conditioner(li, feats)
        for li in range(3):
            dominant(li)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2 * w["K"]
        r0.record()
        for i in range(reps):
            dominant(i % w["K"])          # cycles through the layers: weights stream from HBM/L2 as in a pass
        r1.record()
        torch.cuda.synchronize()
        pass_ms = r0.elapsed_time(r1) / reps
        flops_pass = flops_per_sample_layer(w) * xin.shape[0]
        bf16 = peaks.get("bf16_tflops", 1630.0)
        tensor_peak = bf16 / 2.0
        note = "%s bf16_tflops (burst: kernel timed alone) / 2 (tf32)" % peak_src
        if self.prec == "tf32":
            f_gemm0 = 2.0 * 2 * w["n"] * w["H"]
            f_final = 2.0 * w["H"] * w["n"] * (3 * w["nb"] + 1)
            f_tf32 = f_gemm0 + (0.0 if self.fused else f_final)
            f_f16 = flops_per_sample_layer(w) - f_tf32
            tensor_peak = (f_tf32 + f_f16) / (f_tf32 / (bf16 / 2.0) + f_f16 / bf16)
            note = ("%s bf16_tflops (burst: the kernel is timed alone) for the FP16-operand GEMMs, half of it for the "
                    "TF32 GEMM0, flop-weighted harmonic blend" % peak_src)
        achieved = flops_pass / (pass_ms * 1e-3) / 1e12
        traffic = None
        for tj in ("r02_traffic.json", "r01_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tj)
            if os.path.exists(tp) and self.prec == "tf32":
                # ncu dram bytes per launch of this kernel at this row count (weights dominate)
                traffic = json.load(open(tp)).get("tc_conditioner_kernel<%d>%s@%s@%d" % (
                    w["H"], "+spline" if self.fused else "", self.name, xin.shape[0]))
                if traffic is not None:
                    break
        kern = (("tc_conditioner_kernel (tcgen05 kind::f16 / kind::tf32, fused spline epilogue)" if self.fused else
                 "tc_conditioner_kernel (tcgen05 kind::tf32)") if self.prec == "tf32" else "linear_kernel chain (fp32)")
        one = {"bound": "tensor", "kernel": kern + ", one coupling layer", "achieved": achieved, "peak": tensor_peak,
               "unit": "TFLOP/s", "frac": achieved / tensor_peak, "traffic": traffic, "peak_source": note,
               "launch_ms": pass_ms, "rows": int(xin.shape[0]),
               "algorithmic_flop_per_launch": flops_pass}
        # The flow passes launch the K layers of a pass as ONE grid of K x tiles CTAs when that pays for this flow and
        # pass size (fs_flow_uses_layer_parallel): that launch is then the dominant kernel of the step, the single-layer
        # figure is kept beside it.
        if self.fused and pack.uses_layer_parallel(int(xin.shape[0])):
            try:
                K = w["K"]
                feats_all = feats_t.repeat(K)
                b0, b1 = xin.clone(), torch.zeros_like(xin)
                for _ in range(2):
                    pack.coupling_all("density", feats_all, b0, b1)
                torch.cuda.synchronize()
                reps = 4
                r0.record()
                for _ in range(reps):
                    pack.coupling_all("density", feats_all, b0, b1)
                r1.record()
                torch.cuda.synchronize()
                all_ms = r0.elapsed_time(r1) / reps
                ach = K * flops_pass / (all_ms * 1e-3) / 1e12
                return {"bound": "tensor", "kernel": kern + ", all %d coupling layers of a pass in one launch of %d x %d "
                        "thread blocks (layer-parallel)" % (K, K, (xin.shape[0] + 127) // 128), "achieved": ach,
                        "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                        "traffic": (traffic * K if traffic is not None else None), "peak_source": note,
                        "launch_ms": all_ms, "rows": int(xin.shape[0]), "algorithmic_flop_per_launch": K * flops_pass,
                        "single_layer_launch": one}
            except Exception as ex:        # flows without the layer-parallel path (odd N): one launch per layer
                one["layer_parallel"] = "unavailable: %s" % (str(ex)[:120],)
        return one

    def phases(self, fp32=None, peaks=None):
        """Phase split of one round (not part of the timed region) + the rooflines of the non-dominant kernels."""
        w, eng, model, dev, B, n = self.w, self.eng, self.model, self.dev, self.B, self.n
        out = {}
        if self.trainer is not None:                           # Algorithm-2 cycle: split of one whole cycle
            self.train_cycle()
            torch.cuda.synchronize()
            timers = []
            self.train_cycle(timers=timers)
            torch.cuda.synchronize()
            for (_, a), (name, b_) in zip(timers, timers[1:]):
                out["cycle_" + name + "_ms"] = a.elapsed_time(b_)
            tr = self.trainer
            out["train_steps_per_cycle"] = -(-w["train"]["samples"] // w["train"]["batch"])
            out["gradient_bucket_bytes"] = int(tr.flat.numel() * 4) if tr.flat is not None else 0
            if self.world > 1:
                import torch.distributed as dist
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.all_reduce(tr.flat)
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    dist.all_reduce(tr.flat)
                b_.record()
                torch.cuda.synchronize()
                out["allreduce_ms"] = a.elapsed_time(b_) / 5
                out["allreduce_collective"] = "ncclAllReduce(sum) of the flat float32 gradient bucket, one per optimizer step"
                out["allreduce_busbw_gbs"] = (2 * (self.world - 1) / self.world) * out["gradient_bucket_bytes"] / (out["allreduce_ms"] * 1e-3) / 1e9

        def timed(name, fn, reps=1):
            fn()                                       # untimed first call: workspaces of this shape / stream get allocated
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r = fn()
            b_.record()
            torch.cuda.synchronize()
            out[name] = a.elapsed_time(b_) / reps
            return r
        timed("local_sweep_ms", lambda: eng.particle_displacement(w["local"]))
        zz = model.q0(B)
        cfg = timed("flow_sample_ms", lambda: (model.forward(zz).reshape(B, n, 2) + self.half32).contiguous())
        timed("energy_total_ms", lambda: eng.total_energy_virial(cfg), reps=5)
        xin = eng.centred(torch.cat([eng.pos, eng.pos]))
        lq = timed("flow_log_prob_2B_ms", lambda: model.log_prob(xin))
        E_new, W_new, _ = eng.total_energy_virial(cfg)
        u = torch.rand(B, dtype=torch.float64, device=dev)
        acc0 = int(eng.accepted.sum().item())
        lqo, lqn = lq[:B].contiguous(), lq[B:].contiguous()
        timed("global_move_fused_ms", lambda: eng.nf_big_move(cfg, u=u, logq=(lqo, lqn)), reps=5)
        # the stand-alone acceptance kernel on the same inputs (its HBM figure; the round runs the fused kernel above)
        # 20 launches replayed from a CUDA graph: the kernel takes ~8 us, a ctypes call from Python ~14 us
        mask_buf = torch.empty(B, dtype=torch.uint8, device=dev)
        acc_fn = lambda: eng.accept_global(cfg, E_new, W_new, lqo, lqn, u=u, mask=mask_buf)
        try:
            if self.world > 1:                     # no stream capture next to a live NCCL communicator (its watchdog
                raise RuntimeError("plain launches")   # thread queries events): plain launches there
            acc_fn()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                acc_fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    acc_fn()
            timed("accept_global_ms", gr.replay, reps=3)
            out["accept_global_ms"] /= 20.0
            acc_calls = 2 + 20 * 4
        except Exception:                          # capture refused: plain launches (host-bound figure)
            timed("accept_global_ms", acc_fn, reps=20)
            acc_calls = 21
        extra = {}
        if fp32:
            sw = sweep_flops_per_step(n) * B * w["local"] / (out["local_sweep_ms"] * 1e-3) / 1e12
            extra["local_sweep"] = {"bound": "fp32", "kernel": "local_sweep_fast_kernel", "achieved": sw,
                                    "peak": fp32["peak"], "unit": "TFLOP/s", "frac": sw / fp32["peak"],
                                    "peak_source": "scripts/fp32_peak.cu measured in this run"}
            en = energy_flops(B, n) / (out["energy_total_ms"] * 1e-3) / 1e12
            extra["energy_total"] = {"bound": "fp32", "kernel": "energy_total_kernel_v2", "achieved": en,
                                     "peak": fp32["peak"], "unit": "TFLOP/s", "frac": en / fp32["peak"],
                                     "peak_source": "scripts/fp32_peak.cu measured in this run"}
        if peaks:
            # accept kernel: 8N bytes of proposal read + 8N written per accepted chain + ~40 bytes of scalars per chain
            # (the phase above re-ran the move 6 times on the same inputs: use its acceptance count)
            frac_acc = (int(eng.accepted.sum().item()) - acc0) / ((6.0 + acc_calls) * B)
            bytes_ = B * (8.0 * n * (1 + frac_acc) + 40)
            gbs = bytes_ / (out["accept_global_ms"] * 1e-3) / 1e9
            extra["accept_global"] = {"bound": "hbm", "kernel": "accept_global_kernel", "achieved": gbs,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                      "note": "stand-alone fs_accept_global replayed from a CUDA graph on the same buffers "
                                              "(17 MB of proposals stay in the 126 MB L2 between replays: an L2-assisted "
                                              "figure; cold-cache ncu launch: 8.2 us, DRAM 26 %% of peak, "
                                              "profiles/r02f_ncu_summary.md), %.1f KB in %.1f us per launch; the round runs "
                                              "fs_accept_global_fused"
                                              % (bytes_ / 1e3, out["accept_global_ms"] * 1e3)}
            if fp32:
                gm = energy_flops(B, n) / (out["global_move_fused_ms"] * 1e-3) / 1e12
                extra["global_move_fused"] = {"bound": "fp32", "kernel": "energy_total_kernel_v2<.., ACCEPT> (fs_accept_global_fused)",
                                              "achieved": gm, "peak": fp32["peak"], "unit": "TFLOP/s",
                                              "frac": gm / fp32["peak"],
                                              "peak_source": "scripts/fp32_peak.cu measured in this run"}
        return out, extra


def energy_sweep(dev, fp32, MC):
    """BASELINE configs[4]: fs_energy_total over N = 64 .. 4096 x 1k-64k configurations, against the FP32 peak measured
    in this run.  Algorithmic flop = 27 B N (N-1) / 2 + 40 B N."""
    rows = []
    for n, B in ((64, 65536), (128, 32768), (256, 16384), (512, 8192), (1024, 4096), (2048, 2048), (4096, 1024)):
        pos, L = MC.jittered_lattice(n, 0.5, seed=7, batch=64)
        pos = torch.from_numpy(pos).to(dev).repeat(B // 64, 1, 1).contiguous()
        eng = MC.BatchedMonteCarlo(pos[:64], MC.SimulationBox(L), 1.0, n, rng="philox", device=dev, **POT)
        eng.total_energy_virial(pos)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            eng.total_energy_virial(pos)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tf = energy_flops(B, n) / (ms * 1e-3) / 1e12
        rows.append({"particles": n, "configs": B, "ms": ms, "pairs_per_s": B * n * (n - 1) / 2 / (ms * 1e-3),
                     "achieved": tf, "frac": tf / fp32["peak"] if fp32 else None})
        del pos, eng
    return {"metric": "pair_energy_tflops", "desc": "BASELINE configs[4]: energy-kernel sweep, rho = 0.5 jittered lattices",
            "roofline": {"bound": "fp32", "kernel": "energy_total_kernel_v2", "unit": "TFLOP/s",
                         "peak": fp32["peak"] if fp32 else None, "peak_detail": fp32,
                         "peak_source": "scripts/fp32_peak.cu measured in this run",
                         "achieved": max(r["achieved"] for r in rows),
                         "frac": max(r["frac"] for r in rows) if fp32 else None},
            "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "tf32", "fp32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="override chains per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads / energy sweep")
    ap.add_argument("--budget", type=float, default=6.0, help="seconds of CPU work per reference-arm step")
    ap.add_argument("--parity-file", default=None, help="(reference arm) parity sample dumped by the GPU arm")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.chains:
        w["chains"] = args.chains
    if args.impl == "reference":
        return run_reference(args, args.workload, w)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import flowstate_b200.MCMC as MC

    peaks, peak_src = load_peaks()
    h = Harness(args.workload, w, dev, rank, world, args.precision)
    tmpdir = tempfile.mkdtemp(prefix="fs_parity_")
    parity_path = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity_path = os.path.join(tmpdir, "sample")
        h.dump_parity_sample(parity_path)
    # clock sampler: in-process NVML thread, started (and its first sample awaited) before the warm-up rounds; its
    # samples cover the warm-up rounds (same load) and the timed region
    clocks = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("FS_NO_CLOCKS")) else None
    if clocks:
        clocks.wait_ready()
    ms_total, launches = h.time_device(args.steps, args.warmup)
    main_step_ms = list(h.step_ms)
    main_warmup = h.warmup_rounds
    clk = clocks.stop() if clocks else None
    e2e_s, h2d, d2h = h.time_e2e(args.steps)
    fp32 = fp32_peak(local_rank) if rank == 0 else None
    roof = h.conditioner_roofline(peaks, peak_src)
    phases, extra = h.phases(fp32, peaks)
    variant = None
    if rank == 0 and world == 1 and h.trainer is None and not args.no_secondary:
        # Not the headline: the same round with log q(proposal) taken from the sampling pass that produced the proposal
        # (NormalizingFlow.sample_with_log_prob) instead of a second trip through the density pass - a third of the
        # conditioner work of a round.  The headline keeps the reference's order (monte_carlo.py:262).
        h.pending.clear()
        h.logq_from_sample = True
        v_steps = 12
        try:
            ms_v, l_v = h.time_device(v_steps, 4)
        except Exception as exv:                       # a variant must never take the line down
            ms_v, l_v = float("nan"), 0
            h.step_ms = [float("nan")]
            variant_error = repr(exv)[:200]
        h.logq_from_sample = False
        h.pending.clear()
        variant = {"value": w["chains"] * (w["local"] + 1) * v_steps / (ms_v * 1e-3), "unit": "chain-steps/s",
                   "ms_per_step": ms_v / v_steps, "steps": v_steps, "gpu_launches": l_v,
                   "step_ms_min_median_max": [min(h.step_ms), sorted(h.step_ms)[len(h.step_ms) // 2], max(h.step_ms)],
                   "note": "log q(new) = log q0(z) - log-det of the sampling pass; parity: "
                           "tests/test_gpu_flow.py::test_sampling_pass_log_prob_equals_inverse_pass, "
                           "tests/test_gpu_global.py::test_global_move_with_sampling_pass_log_density"}

    graphed = None
    if rank == 0 and world == 1 and h.trainer is None and not args.no_secondary:
        graphed = time_graphed_round(h, 20)
    secondary = {}
    if rank == 0 and world == 1 and not args.no_secondary:
        del h.model, h.eng
        h.pending.clear()
        torch.cuda.empty_cache()
        for name in WORKLOADS:
            if name == args.workload:
                continue
            try:
                ws = dict(WORKLOADS[name])
                hs = Harness(name, ws, dev, 0, 1, args.precision)
                # short rounds: enough of them that the pipeline refill after the bracketing synchronize (about four
                # rounds, see time_device) weighs as little as in the headline run
                s_steps = 12 if (ws.get("train") or ws["n"] >= 128) else 60
                ms_s, l_s = hs.time_device(s_steps, 4)
                e2e_ss, h2d_s, d2h_s = hs.time_e2e(s_steps)
                tot = ws["chains"] * (ws["local"] + 1) * s_steps
                ph, ex = hs.phases(fp32, peaks)
                gr_s = time_graphed_round(hs, s_steps) if hs.trainer is None else None
                secondary[name] = {"metric": "mh_chain_steps_per_s", "value": tot / (ms_s * 1e-3), "unit": "chain-steps/s",
                                   "steps": s_steps, "warmup": hs.warmup_rounds, "ms_per_step": ms_s / s_steps, "gpu_launches": l_s,
                                   "step_ms_min_median_max": [min(hs.step_ms), sorted(hs.step_ms)[len(hs.step_ms) // 2],
                                                              max(hs.step_ms)],
                                   "config": workload_config(name, ws),
                                   "e2e": {"value": tot / e2e_ss, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d_s,
                                           "d2h_bytes_per_step": d2h_s},
                                   "roofline": hs.conditioner_roofline(peaks, peak_src), "phases_ms": ph,
                                   "other_rooflines": ex, "variant_graphed_round": gr_s}
                del hs
                torch.cuda.empty_cache()
            except Exception as ex:      # a secondary entry must never take the line down
                secondary[name] = {"error": repr(ex)[:300]}
                torch.cuda.empty_cache()
        try:
            secondary["energy_sweep"] = energy_sweep(dev, fp32, MC)
        except Exception as ex:      # a secondary entry must never take the line down
            secondary["energy_sweep"] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n, B = w["n"], w["chains"]
    steps_total = world * B * (w["local"] + 1) * args.steps
    value = steps_total / (ms_total * 1e-3)
    line = {
        "metric": "mh_chain_steps_per_s", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if h.prec == "tf32" else "f32", "data": "synthetic",
        "config": workload_config(args.workload, w),
        "details": {"rng": "philox",
                    "conditioner": ("tf32 operands in GEMM0, fp16 operands in the residual blocks and the final layer, "
                                    "fp32 accumulation" if h.fused else h.prec),
                    "pipelining": "proposals of round r+1 sampled on a side stream during round r",
                    "weight_broadcast_bytes": h.bcast_bytes,
                    "timed_region_ms": ms_total,
                    "warmup_rounds_run": main_warmup,
                    "variant_logq_from_sampling_pass": variant,
                    "variant_graphed_round": graphed,
                    "step_ms_min_median_max": [min(main_step_ms), sorted(main_step_ms)[len(main_step_ms) // 2],
                                               max(main_step_ms)]},
        "nf_proposals_per_s": world * B * args.steps / (ms_total * 1e-3),
        "gpu_launches": launches, "clocks": clk, "phases_ms": phases,
        "e2e": {"value": steps_total / e2e_s, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "roofline": roof, "other_rooflines": extra, "fp32_peak_tflops": fp32,
    }
    if secondary:
        line["secondary"] = secondary
    if not args.no_cpu_baseline and world == 1:
        # separate process: the oracle forks worker processes, which must not inherit a CUDA context
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--workload", args.workload, "--budget", "20"]
            if args.chains:
                cmd += ["--chains", str(args.chains)]
            if parity_path:
                cmd += ["--parity-file", parity_path]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
                env.pop(k, None)
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
            base = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
            par = base.pop("parity", None)
            line["cpu_baseline"] = base
            if par is not None:
                line["parity_checked"] = bool(par.get("ok"))
                line["parity"] = par
        except Exception as ex:     # the baseline must never take the GPU line down
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
