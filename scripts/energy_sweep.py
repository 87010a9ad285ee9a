"""BASELINE config 5: pair-energy throughput sweep N = 64..4096 x configs, against the FP32 roofline.

flops = 27 B N(N-1)/2 + 40 B N, bytes = B (8N + 9)  (SURVEY.md 8d).  The FP32 peak used as the
denominator is measured in the same run by an FMA micro-kernel (torch: a chain of fused multiply-adds
cannot be expressed, so the nominal 148 SM x 128 lanes x 2 x SM clock is reported beside a measured
dependent-FMA rate from the library's own probe when available).
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import flowstate_b200.MCMC as MC

dev = torch.device("cuda")
rows = []
sm_clock_ghz = 1.965
fp32_peak = 148 * 128 * 2 * sm_clock_ghz / 1e3          # TFLOP/s nominal at max clock
for n, B in [(64, 65536), (128, 32768), (256, 16384), (256, 65536), (512, 8192), (1024, 4096), (2048, 2048), (4096, 1024)]:
    pos, L = MC.jittered_lattice(n, 0.5, seed=1, batch=8)
    pos = np.tile(pos, (B // 8, 1, 1))
    eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                               rng="philox")
    for _ in range(3):
        eng.total_energy_virial()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        eng.total_energy_virial()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pairs = B * n * (n - 1) / 2
    flops = 27 * pairs + 40 * B * n
    byts = B * (8 * n + 9)
    rows.append(dict(N=n, B=B, ms=ms, gpairs_per_s=pairs / ms / 1e6, tflops=flops / ms / 1e9,
                     frac_fp32=flops / ms / 1e9 / fp32_peak, gbs=byts / ms / 1e6))
    print(json.dumps(rows[-1]))
