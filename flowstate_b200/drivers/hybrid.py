"""Hybrid NF-MCMC drivers on the batched GPU engine.

The reference's drivers are module-level scripts with constants in the source
(hybrid_NF_MCMC/main_algorithm_1.py:33-73, main_algorithm_2.py:33-76) that loop over
MonteCarlo objects in Python.  These functions run the same algorithms with every chain
resident on the GPU (BatchedMonteCarlo), take the constants as a config object / CLI
flags, and carry the multi-GPU hooks (chains sharded by rank, weights broadcast after
training, gradients all-reduced in Algorithm 2).  Defaults are the reference's constants.

    python -m flowstate_b200.drivers.hybrid --algorithm 0 --particles 3 --chains 100     # MCMC only
    python -m flowstate_b200.drivers.hybrid --algorithm 1 --particles 3 --chains 10
    torchrun --nproc-per-node 8 -m flowstate_b200.drivers.hybrid --algorithm 2 --particles 64 --chains 4096

MCMC only (main_mcmc_only.py:176-236): equilibrate, then `production_steps` local moves per chain sampled every
`sampling_frequency` steps; per-chain well statistics / free-energy difference from the sampled configurations.
Algorithm 1 (main_algorithm_1.py:203-395): equilibrate -> collect local samples -> train the
flow by forward KL -> interleave `big_move_interval` local moves with one NF global move.
Algorithm 2 (main_algorithm_2.py:393-577): per cycle `local_steps` local moves, a short
training pass on the newest samples, one NF global move per chain.  The loss is
alpha * forward KL + (1 - alpha) * reverse KL against NF.Energy.DoubleWellLJ; with the
reference's ALPHA = 1.0 (main_algorithm_2.py:52) the reverse-KL term has weight 0 and is
not evaluated.
"""
import argparse
import dataclasses
import json
import time

import numpy as np
import torch
import torch.distributed as dist

from .. import MCMC as MC
from .. import normflows as NF
from .. import parallel


@dataclasses.dataclass
class HybridConfig:
    """Defaults = Algorithm 1's constants (main_algorithm_1.py:33-73); HybridConfig.preset(algorithm) applies the other
    drivers' constants (main_algorithm_2.py:33-76, main_mcmc_only.py:40-60) on top."""
    particles: int = 3
    chains: int = 10                 # total over all ranks (NUM_MC_RUNS)
    rho: float = 0.03
    temperature: float = 1.0
    V0_list: tuple = (-10.0, -10.5)
    r0: float = 1.2
    k: float = 15.0
    max_displacement: float = 0.65
    equilibration_steps: int = 5000
    adjusting_frequency: int = 5000
    sampling_frequency: int = 150
    master_seed: int = 42
    # random streams of the chains: "philox" = counter-based streams keyed by the global chain id (the throughput kernel,
    # results independent of the number of GPUs); "pcg64" = numpy-compatible per-chain generators seeded
    # master_seed + chain id like the reference (monte_carlo.py:92-95; the replay / parity kernel)
    rng: str = "philox"
    # flow (Alg-1: K=15, H=256, NUM_BINS passed as num_blocks -> 32 blocks, 32 bins; main_algorithm_1.py:63-70, 282)
    K: int = 15
    blocks: int = 32
    hidden: int = 256
    bins: int = 32
    lr: float = 1e-4
    weight_decay: float = 0.0
    batch_size: int = 512
    epochs: int = 100
    training_samples: int = 102400
    # hybrid phase
    big_move_attempts: int = 1000
    big_move_interval: int = 1000
    # Alg 2
    cycles: int = 1000
    local_steps: int = 100
    alpha: float = 1.0               # loss = alpha * forward KL + (1 - alpha) * reverse KL (main_algorithm_2.py:52, 321)
    precision: str = "auto"          # conditioner arithmetic of the eval-mode kernels: auto | tf32 | fp32
    cuda_graph: int = 1              # replay the training pass from a CUDA graph (0: eager autograd)
    sync_bn: int = 0                 # N > 1 GPUs: BatchNorm batch statistics over all ranks' batches (SyncBatchNorm)
    # MCMC only (main_mcmc_only.py:56-57: 1e7 steps over 100 chains)
    production_steps: int = 100000

    @staticmethod
    def preset(algorithm, **overrides):
        cfg = HybridConfig(**PRESETS[int(algorithm)])
        return dataclasses.replace(cfg, **overrides)


PRESETS = {
    # main_mcmc_only.py:33-60
    0: dict(chains=100, production_steps=100000),
    # main_algorithm_1.py:33-73
    1: dict(),
    # main_algorithm_2.py:33-76: 100 chains, K=23, H=128, 2 blocks, 15 bins, Adam lr 5.435e-4 wd 9.586e-5, batch 256,
    # one epoch per cycle on the newest 1000 samples (100 local steps per chain sampled every 10), no adaptation during
    # the cycles (ADJUSTING_FREQUENCY 10000 > EQUILIBRATION_STEPS 5000)
    2: dict(chains=100, K=23, blocks=2, hidden=128, bins=15, lr=0.000543510751759681,
            weight_decay=9.5857178422352e-05, batch_size=256, epochs=1, training_samples=1000, sampling_frequency=10,
            adjusting_frequency=10000, cycles=1000, local_steps=100),
}


def _dist():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _init_chains(cfg, device):
    # Everything allocated so far (imports, modules) goes to the collector's permanent generation: a full collection of
    # this heap in the middle of a run stalls kernel submission for ~0.1 s (bench.py: quiet_host).
    import gc
    gc.collect()
    gc.freeze()
    rank, world = _dist()
    start, count = parallel.shard_range(cfg.chains, rank, world)
    # even chains start in the left well, odd chains in the right one (main_algorithm_1.py:149-165); the box length is
    # rounded to float32 so the device and any float64 re-evaluation see the same box
    pos, _ = MC.initialise_chains(count, cfg.particles, cfg.rho, first_chain=start)
    L = float(np.float32(np.sqrt(cfg.particles / cfg.rho)))
    box = MC.SimulationBox(L)
    kw = dict(num_wells=2, V0_list=list(cfg.V0_list), r0=cfg.r0, k=cfg.k,
              initial_max_displacement=cfg.max_displacement, device=device)
    if cfg.rng == "pcg64":
        kw.update(seeds=[cfg.master_seed + start + i for i in range(count)])
    else:
        kw.update(rng="philox", philox_seed=cfg.master_seed, chain_id0=start)
    eng = MC.BatchedMonteCarlo(pos, box, cfg.temperature, cfg.particles, **kw)
    return eng, L


def _build_flow(cfg, L, device):
    bound = L / 2
    base = NF.Energy.UniformParticle(cfg.particles, 2, bound, device=device)
    D = 2 * cfg.particles
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(D, cfg.blocks, cfg.hidden, range(D), num_bins=cfg.bins,
                                                              tail_bound=bound) for _ in range(cfg.K)]
    target = NF.Energy.DoubleWellLJ(D, cfg.particles, cfg.temperature, bound, V0_list=list(cfg.V0_list), r0=cfg.r0,
                                    k=cfg.k)                                     # main_algorithm_2.py:282-285
    model = NF.NormalizingFlow(base, layers, target).to(device)
    model.precision = cfg.precision
    # every rank starts from rank 0's weights (gradients are averaged, so the replicas then stay identical); the
    # proposal / permutation generator is seeded per rank so chains on different GPUs see different base noise
    parallel.broadcast_flow(model, src=0)
    rank, _ = _dist()
    torch.manual_seed(cfg.master_seed + 7919 * (rank + 1))
    return model


def _local_phase(eng, steps, cfg, collect=None, step0=0):
    """`steps` local moves per chain, adapting / sampling at the reference's frequencies."""
    done = 0
    while done < steps:
        nxt = steps
        for f in (cfg.adjusting_frequency, cfg.sampling_frequency if collect is not None else 0):
            if f:
                nxt = min(nxt, ((step0 + done) // f + 1) * f - step0)
        n = max(1, min(nxt, steps) - done)
        eng.particle_displacement(n)
        done += n
        g = step0 + done
        if cfg.adjusting_frequency and g % cfg.adjusting_frequency == 0:
            eng.adjust_displacement()
        if collect is not None and g % cfg.sampling_frequency == 0:
            collect.append(eng.centred(eng.pos).clone())
    return step0 + done


def _train(model, data, cfg, epochs, trainer=None):
    """Training (main_algorithm_1.py:297-320, main_algorithm_2.py:437-452) through drivers.training.FlowTrainer: forward +
    backward replayed from a CUDA graph, ONE all-reduce of the flat gradient bucket per optimizer step, fused Adam, a new
    optimizer per call (the reference creates a new Adam every cycle, main_algorithm_2.py:440).  The number of
    collectives is the same on every rank: the number of minibatches per epoch is the minimum over ranks (ranks may own
    different numbers of chains), and the reference's "skip a NaN / Inf loss" decision (main_algorithm_1.py:310-315) is
    taken collectively.  Ends in eval mode with rank 0's weights and BatchNorm statistics on every rank."""
    from .training import FlowTrainer
    model.train()
    tr = trainer or FlowTrainer(model, cfg.lr, cfg.weight_decay, cfg.alpha, cfg.batch_size, use_graph=cfg.cuda_graph,
                                   sync_bn=cfg.sync_bn)
    tr.fresh_optimizer()
    _, world = _dist()
    n_batches = torch.tensor([max(0, -(-(data.shape[0] - 1) // cfg.batch_size))], device=data.device)
    if world > 1:
        dist.all_reduce(n_batches, op=dist.ReduceOp.MIN)
    n_batches = int(n_batches.item())
    losses = []
    for _ in range(epochs):
        perm = torch.randperm(data.shape[0], device=data.device)
        tot, nb = 0.0, 0
        dev_losses = []
        for bi in range(n_batches):
            batch = data[perm[bi * cfg.batch_size:(bi + 1) * cfg.batch_size]]
            if tr.native_adam:                                 # no host wait per step: losses are read once per epoch
                dev_losses.append(tr.step(batch, sync=False))
                continue
            loss = tr.step(batch)
            if loss is not None:
                tot += loss
                nb += 1
        if dev_losses:
            ls = torch.cat([l.reshape(1) for l in dev_losses])
            ok = torch.isfinite(ls)                            # a skipped step shows as a non-finite loss
            tot, nb = float(ls[ok].sum()), int(ok.sum())
        losses.append(tot / max(nb, 1))
    model.eval()
    if not tr.sync_bn:            # parameters followed the all-reduced gradients on every rank; the BatchNorm running
        parallel.broadcast_flow(model, src=0, buffers_only=True)   # statistics are per rank unless synchronised
    return losses


def run_mcmc_only(cfg, device="cuda", log=print):
    """main_mcmc_only.py:176-236 on the batched engine: equilibration, production with configurations sampled every
    `sampling_frequency` steps, then calculate_well_statistics per chain (cumulative p_A, p_B, dF = ln(p_B / p_A))."""
    from . import observables
    eng, L = _init_chains(cfg, device)
    step = _local_phase(eng, cfg.equilibration_steps, cfg)
    samples = []                                             # centred coordinates, one [B, 2N] tensor per sample time
    cfg_prod = dataclasses.replace(cfg, adjusting_frequency=0)   # the reference adapts during equilibration only
    _local_phase(eng, cfg.production_steps, cfg_prod, collect=samples, step0=step)
    B, n = eng.B, cfg.particles
    half_box = L / 2
    traj = torch.stack(samples, dim=1).reshape(B, len(samples), n, 2) + np.float32(half_box)   # MC-box coordinates
    delta_f, p_a, p_b = [], [], []
    _, state, _ = observables._classify(traj.reshape(B * len(samples), n, 2), half_box, cfg.r0)
    state = state.reshape(B, len(samples)).cpu().numpy()
    runs = np.arange(1, len(samples) + 1)
    for b in range(B):
        pa = np.cumsum(state[b] == 1) / runs
        pb = np.cumsum(state[b] == 2) / runs
        p_a.append(float(pa[-1]))
        p_b.append(float(pb[-1]))
        delta_f.append(float(np.log(pb[-1] / pa[-1])) if pa[-1] > 0 and pb[-1] > 0 else 0.0)
    att, acc, _ = parallel.allreduce_counters(eng.attempts, eng.accepted, torch.zeros(1, device=eng.device))
    log("%d chains x %d production steps, acceptance %.3f, mean dF %.3f" % (B, cfg.production_steps,
                                                                         acc / max(att, 1), float(np.mean(delta_f))))
    return {"algorithm": 0, "attempts": att, "accepted": acc, "p_a": p_a, "p_b": p_b, "delta_f": delta_f,
            "samples_per_chain": len(samples), "engine": eng}


def run_algorithm_1(cfg, device="cuda", log=print):
    eng, L = _init_chains(cfg, device)
    step = _local_phase(eng, cfg.equilibration_steps, cfg)
    samples = []
    per_chain = max(1, -(-cfg.training_samples // cfg.chains))
    cfg_prod = dataclasses.replace(cfg, adjusting_frequency=0)   # the reference adapts during equilibration only
    _local_phase(eng, per_chain * cfg.sampling_frequency, cfg_prod, collect=samples, step0=step)   # (main_algorithm_1.py:203-210 vs 245-252)
    data = torch.cat(samples, dim=0)
    model = _build_flow(cfg, L, device)
    t0 = time.time()
    losses = _train(model, data, cfg, cfg.epochs)
    log("trained %d epochs on %d samples in %.1fs, loss %.4f -> %.4f" % (cfg.epochs, data.shape[0], time.time() - t0,
                                                                         losses[0], losses[-1]))
    eng.set_nf_model(model)
    big_acc = 0
    for _ in range(cfg.big_move_attempts):
        eng.particle_displacement(cfg.big_move_interval)
        big_acc += int(eng.nf_big_move().sum().item())
    att, acc, big = parallel.allreduce_counters(eng.attempts, eng.accepted, torch.tensor([big_acc], device=eng.device))
    return {"algorithm": 1, "attempts": att, "accepted": acc, "big_move_accepts": big,
            "big_move_attempts": cfg.big_move_attempts * cfg.chains, "final_loss": losses[-1], "engine": eng,
            "model": model}


def _collect(eng, steps, cfg):
    """`steps` local moves per chain, configurations (centred coordinates) sampled every cfg.sampling_frequency."""
    samples = []
    sf = max(1, cfg.sampling_frequency)
    for s in range(0, steps, sf):
        n = min(sf, steps - s)
        eng.particle_displacement(n)
        if n == sf:
            samples.append(eng.centred(eng.pos).clone())
    return torch.cat(samples, dim=0) if samples else eng.centred(eng.pos).clone()


def run_algorithm_2(cfg, device="cuda", log=print):
    """main_algorithm_2.py:209-577: equilibrate, initial training on `training_samples` configurations, then per cycle
    `local_steps` local moves per chain (sampled every `sampling_frequency`), `epochs` of training on the new samples
    with a fresh Adam (main_algorithm_2.py:440), one NF global move per chain."""
    eng, L = _init_chains(cfg, device)
    _local_phase(eng, cfg.equilibration_steps, cfg)
    model = _build_flow(cfg, L, device)
    model.layer_parallel = "prefer"        # every sample / log_prob pass of a cycle runs alone on the GPU
    eng.set_nf_model(model)
    # initial training set: INITIAL_TRAINING_NUM_SAMPLES / (NUM_MC_RUNS / SAMPLING_FREQUENCY) steps per chain
    # (main_algorithm_2.py:240-252)
    from .training import FlowTrainer
    trainer = FlowTrainer(model, cfg.lr, cfg.weight_decay, cfg.alpha, cfg.batch_size, use_graph=cfg.cuda_graph,
                                   sync_bn=cfg.sync_bn)
    init_steps = max(cfg.sampling_frequency, int(cfg.training_samples / (cfg.chains / cfg.sampling_frequency)))
    last = _train(model, _collect(eng, init_steps, cfg), cfg, cfg.epochs, trainer)[-1]
    big_acc = 0
    for cycle in range(cfg.cycles):
        data = _collect(eng, cfg.local_steps, cfg)
        last = _train(model, data, cfg, cfg.epochs, trainer)[-1]          # graphs are reused, the optimizer is new
        big_acc += int(eng.nf_big_move().sum().item())
    att, acc, big = parallel.allreduce_counters(eng.attempts, eng.accepted, torch.tensor([big_acc], device=eng.device))
    return {"algorithm": 2, "attempts": att, "accepted": acc, "big_move_accepts": big,
            "big_move_attempts": cfg.cycles * cfg.chains, "final_loss": last, "engine": eng, "model": model}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--algorithm", type=int, choices=(0, 1, 2), default=1, help="0 = local-displacement MCMC only")
    scalar = [f for f in dataclasses.fields(HybridConfig) if f.type in (int, float, str)]
    for f in scalar:                                          # default None: taken from the algorithm's preset
        ap.add_argument("--" + f.name.replace("_", "-"), type=f.type, default=None)
    a = ap.parse_args(argv)
    cfg = HybridConfig.preset(a.algorithm, **{f.name: getattr(a, f.name) for f in scalar
                                              if getattr(a, f.name) is not None})
    import os
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    run = {0: run_mcmc_only, 1: run_algorithm_1, 2: run_algorithm_2}[a.algorithm]
    out = run(cfg, device="cuda:%d" % local)
    if _dist()[0] == 0:
        print(json.dumps({k: v for k, v in out.items() if k not in ("engine", "model")}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
