#!/usr/bin/env python
"""Benchmark of the NF-MCMC sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload alg1_n32|alg1_n256|alg2_n64]
                    [--precision auto|tf32|fp32] [--impl ours|reference]

One "step" = one hybrid Algorithm-1 round on every chain of this rank
(hybrid_NF_MCMC/main_algorithm_1.py:381-395): BIG_MOVE_INTERVAL local
displacement moves + ONE NF-proposed global move (flow sample, total energy of
the proposal, log q of old and new state, Metropolis accept).  The metric is MH
chain-steps/s (every local move and every global move is one chain-step, like
the reference's attempts counter); NF proposals + energy evals/s is reported
beside it.  Default workload: BASELINE configs[1] (4096 chains per GPU, N=32,
Alg-1 flow K=15/H=256/32 blocks/32 bins); chains shard over ranks with no
data-path collective (weak scaling), flow weights are broadcast once from rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: particles, chains per GPU, local steps per round, flow (K, blocks, H, bins), sigma, rho
    "alg1_n32": dict(n=32, chains=4096, local=1000, K=15, blocks=32, H=256, nb=32, sigma=0.02, rho=0.03,
                     desc="BASELINE configs[1]: Alg 1 hybrid, 4096 chains/GPU, N=32"),
    "alg1_n256": dict(n=256, chains=8192, local=1000, K=15, blocks=32, H=256, nb=32, sigma=0.02, rho=0.03,
                      desc="BASELINE configs[2]: Alg 1 hybrid, N=256, 8192 chains/GPU (65536 over 8 GPUs)"),
    "alg2_n64": dict(n=64, chains=4096, local=100, K=23, blocks=2, H=128, nb=15, sigma=0.05, rho=0.03,
                     desc="BASELINE configs[3] sampling part: Alg 2 cycle, N=64, 100 local steps + 1 global move"),
}
POT = dict(num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15)


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


# ---------------------------------------------------------------------------
def run_reference(args, w):
    """--impl reference: the reference's CPU path.  The reference is pure Python and does not
    travel to the GPU box (nothing to compile into oracle/_ref), so this arm times the oracle
    port of it on all host cores, on the same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import flowstate_b200.normflows as NF
    bound = float(np.float32(np.sqrt(w["n"] / w["rho"]))) / 2
    model = build_flow(NF, w, bound, "cpu").eval()
    sd = {k: v for k, v in model.state_dict().items()}
    vals = []
    base = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(w, budget_s=args.budget, state_dict=sd, bound=bound)
        if i >= args.warmup:
            vals.append(base["value"])
        if time.perf_counter() - t_all > 240:
            break
    v = float(np.mean(vals)) if vals else base["value"]
    base["value"] = v
    line = {"impl": "reference", "metric": "mh_chain_steps_per_s", "value": v, "unit": "chain-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (w["local"] + 1) * w["chains"] * args.gpus / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"], "chains_per_gpu": w["chains"],
                       "particles": w["n"], "local_steps_per_round": w["local"]},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--workload", default="alg1_n32", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="auto", choices=["auto", "tf32", "fp32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="override chains per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--budget", type=float, default=6.0, help="seconds of CPU work per reference-arm step")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.chains:
        w["chains"] = args.chains
    if args.impl == "reference":
        return run_reference(args, w)
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
    import flowstate_b200.normflows as NF
    from flowstate_b200 import _lib, parallel

    n, B = w["n"], w["chains"]
    L = float(np.float32(np.sqrt(n / w["rho"])))
    bound = L / 2
    # flow: built on every rank from the same seed on CPU, then broadcast from rank 0 over NCCL
    model = build_flow(NF, w, bound, dev).to(dev).eval()
    bcast_bytes = parallel.broadcast_flow(model, src=0) if world > 1 else 0
    prec = args.precision
    if prec == "auto":
        prec = "tf32"
        try:
            model.precision = "tf32"
            model.log_prob(torch.zeros(2, 2 * n, device=dev))
        except Exception:
            prec = "fp32"
    model.precision = prec
    # chains: this rank's contiguous block of global chain ids
    start, _ = parallel.shard_range(B * world, rank, world)
    pos, _ = MC.jittered_lattice(n, w["rho"], seed=1000 + start, batch=B)
    eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, initial_max_displacement=0.65, rng="philox",
                               philox_seed=20261018, chain_id0=start, device=dev, **POT)
    eng.set_nf_model(model)
    half32 = np.float32(bound)
    torch.manual_seed(1234 + rank)

    # Proposals do not depend on the chains (Alg 1 even pre-generates its whole pool,
    # main_algorithm_1.py:340-343), so the sampling pass of round r+1 runs on a side stream in the shadow
    # of round r's sweep + log-density pass; every round still does one sample pass, two log-densities,
    # one proposal energy and `local` local moves per chain.
    side = torch.cuda.Stream(device=dev)
    pending = {}

    def launch_proposals(z_host=None):
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(side):
            if z_host is None:
                z = model.q0(B)
            else:
                z = torch.empty(B, 2 * n, dtype=torch.float32, device=dev)
                z.copy_(z_host, non_blocking=True)
            cfg = (model.forward(z).reshape(B, n, 2) + half32).contiguous()
            ev = torch.cuda.Event()
            ev.record(side)
        cfg.record_stream(main)
        pending["cfg"], pending["ev"] = cfg, ev

    def one_round(z_host=None):
        if "cfg" not in pending:
            launch_proposals(z_host)
        cfg, ev = pending.pop("cfg"), pending.pop("ev")
        launch_proposals(z_host)                       # next round's proposals, off the critical path
        eng.particle_displacement(w["local"])
        torch.cuda.current_stream(dev).wait_event(ev)
        return eng.nf_big_move(cfg)

    # ---- device-resident timing ------------------------------------------------
    # clock sampler: in-process NVML thread, started (and its first sample awaited) before the warm-up rounds; its
    # samples cover the warm-up rounds (same load) and the timed region
    clocks = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("FS_NO_CLOCKS")) else None
    if clocks:
        clocks.wait_ready()
    if world > 1:
        dist.barrier()
    for _ in range(args.warmup):
        one_round()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.lib().fs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        one_round()
    e1.record()
    torch.cuda.synchronize()
    launches = int(_lib.lib().fs_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.stop() if clocks else None

    # ---- end to end: host buffers in, host results out, copies inside the timed region ----
    # Every step uploads that step's chain positions and base noise from pinned host memory, runs the round and
    # reads positions, energies and accept mask back.  Two host-resident chain batches alternate (batch k % 2 in
    # step k, its output becoming its next input), so the host waits for the results of step k-1 while step k is
    # already queued: the copies and the launch work of consecutive steps overlap as in a real pipeline.
    h_z = torch.empty(B, 2 * n, dtype=torch.float32).uniform_(-bound, bound).pin_memory()
    hb = []
    for k in range(2):
        hp = torch.empty(B, n, 2, dtype=torch.float32).pin_memory()
        hp.copy_(eng.pos.cpu())
        hb.append({"pos": hp, "out_pos": torch.empty(B, n, 2, dtype=torch.float32).pin_memory(),
                   "out_E": torch.empty(B, dtype=torch.float64).pin_memory(),
                   "out_mask": torch.empty(B, dtype=torch.uint8).pin_memory(), "ev": None})

    def e2e_round(k):
        buf = hb[k % 2]
        if buf["ev"] is not None:                      # results of this batch's previous step have landed
            buf["ev"].synchronize()
            buf["pos"].copy_(buf["out_pos"])
        eng.pos.copy_(buf["pos"], non_blocking=True)
        eng.refresh_energy()
        mask = one_round(h_z)                          # base noise comes from the host buffer
        buf["out_pos"].copy_(eng.pos, non_blocking=True)
        buf["out_E"].copy_(eng.E, non_blocking=True)
        buf["out_mask"].copy_(mask, non_blocking=True)
        buf["ev"] = torch.cuda.Event()
        buf["ev"].record()

    for k in range(2):
        e2e_round(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_round(k)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h2d = B * n * 2 * 4 + B * 2 * n * 4
    d2h = B * n * 2 * 4 + B * 8 + B

    # ---- roofline of the dominant kernel: the conditioner of one coupling layer (one launch per layer and pass),
    #      timed alone with CUDA events on the launching stream, on the row count of the log-density pass ----
    xin = eng.centred(torch.cat([eng.pos, eng.pos]))
    model.log_prob(xin)
    pack = model._cuda_pack()
    feats = torch.cat([torch.cos(xin[:, 0::2] * (np.pi / bound)), torch.sin(xin[:, 0::2] * (np.pi / bound))], dim=1).contiguous()
    # the kernel the passes launch: conditioner with the fused spline epilogue when the flow shape has it
    # (H = 128 / 256, at most 32 bins), else the conditioner writing theta
    fused = prec == "tf32" and w["H"] in (128, 256) and w["nb"] <= 32 and not os.environ.get("FS_NO_FUSE")
    xo, ldo = torch.zeros_like(xin), torch.zeros(xin.shape[0], device=dev)

    def dominant(li):
        if fused:
            pack.coupling(li, "density", feats, xin, xo, ldo)
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
    # phase split of one round (not part of the timed region above)
    phases = {}

    def timed(name, fn):
        fn()                                           # untimed first call: workspaces of this shape / stream get allocated
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b_.record()
        torch.cuda.synchronize()
        phases[name] = a.elapsed_time(b_)
        return out
    timed("local_sweep_ms", lambda: eng.particle_displacement(w["local"]))
    zz = model.q0(B)
    cfg = timed("flow_sample_ms", lambda: (model.forward(zz).reshape(B, n, 2) + half32).contiguous())
    timed("energy_total_ms", lambda: eng.total_energy_virial(cfg))
    timed("flow_log_prob_2B_ms", lambda: model.log_prob(xin))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured" if peaks else "fallback"
    # TF32 tensor peak = half the measured BF16 figure (same pipe, K=8 instead of 16 per instruction)
    tensor_peak = bf16 / 2.0
    peak_note = "%s bf16_tflops_sustained / 2 (tf32)" % peak_src
    if prec == "tf32":
        # tensor path: GEMM0 (features) has TF32 operands, every other GEMM FP16 operands (BF16-rate pipe); on the
        # theta path (no fused epilogue) the final layer is TF32 too.  Roofline of the launch = flop-weighted harmonic
        # blend of the two peaks.
        f_gemm0 = 2.0 * 2 * w["n"] * w["H"]
        f_final = 2.0 * w["H"] * w["n"] * (3 * w["nb"] + 1)
        f_tf32 = f_gemm0 + (0.0 if fused else f_final)
        f_f16 = flops_per_sample_layer(w) - f_tf32
        tensor_peak = (f_tf32 + f_f16) / (f_tf32 / (bf16 / 2.0) + f_f16 / bf16)
        peak_note = ("%s bf16_tflops_sustained for the FP16-operand GEMMs, half of it for the TF32 ones, flop-weighted "
                     "harmonic blend" % peak_src)
    achieved = flops_pass / (pass_ms * 1e-3) / 1e12
    traffic = None
    tj = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tj) and prec == "tf32":
        # ncu dram bytes per launch of this kernel at this row count (weights dominate)
        traffic = json.load(open(tj)).get("tc_conditioner_kernel<%d>%s@%s@%d" % (w["H"], "+spline" if fused else "",
                                                                                 args.workload, xin.shape[0]))
    steps_total = world * B * (w["local"] + 1) * args.steps
    value = steps_total / (ms_total * 1e-3)
    line = {
        "metric": "mh_chain_steps_per_s", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if prec == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "chains_per_gpu": B, "particles": n,
                   "local_steps_per_round": w["local"],
                   "flow": {"K": w["K"], "blocks": w["blocks"], "H": w["H"], "bins": w["nb"], "sigma": w["sigma"]},
                   "rho": w["rho"], "rng": "philox",
                   "conditioner": ("tf32 operands in GEMM0, fp16 operands in the residual blocks and the final layer, fp32 "
                                   "accumulation" if fused else prec),
                   "pipelining": "proposals of round r+1 sampled on a side stream during round r",
                   "l2": "inputs larger than L2: %.0f MB of flow weights streamed per pass"
                         % (sum(p.numel() for p in model.parameters()) * 4 / 1e6),
                   "weight_broadcast_bytes": bcast_bytes},
        "nf_proposals_per_s": world * B * args.steps / (ms_total * 1e-3),
        "gpu_launches": launches, "clocks": clk, "phases_ms": phases,
        "e2e": {"value": steps_total / e2e_s, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "roofline": {"bound": "tensor", "kernel": (("tc_conditioner_kernel (tcgen05 kind::f16 / kind::tf32, fused spline epilogue)" if fused else
                                 "tc_conditioner_kernel (tcgen05 kind::tf32)") if prec == "tf32"
                                else "linear_kernel chain (fp32)") + ", one coupling layer",
                     "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                     "traffic": traffic, "peak_source": peak_note,
                     "launch_ms": pass_ms, "rows": int(xin.shape[0])},
    }
    if not args.no_cpu_baseline and world == 1:
        # separate process: the oracle forks worker processes, which must not inherit a CUDA context
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--workload", args.workload, "--budget", "20"]
            if args.chains:
                cmd += ["--chains", str(args.chains)]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
                env.pop(k, None)
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            line["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:     # the baseline must never take the GPU line down
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
