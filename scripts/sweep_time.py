"""Times fs_local_sweep (Philox throughput kernel) at the bench shapes for each lane-group size.
    python scripts/sweep_time.py            # runs itself once per FS_SWEEP_LPC in {8, 16, 32}
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import flowstate_b200.MCMC as MC
    out = {}
    for name, n, B, rho in (("alg1_n32", 32, 4096, 0.03), ("alg1_n256", 256, 8192, 0.03), ("alg2_n64", 64, 4096, 0.03),
                            ("n32_rho0.5", 32, 4096, 0.5), ("n3", 3, 4096, 0.03)):
        L = float(np.float32(np.sqrt(n / rho)))
        pos, _ = MC.jittered_lattice(n, rho, seed=1, batch=B)
        eng = MC.BatchedMonteCarlo(pos, MC.SimulationBox(L), 1.0, n, num_wells=2, V0_list=[-10.0, -10.5], r0=1.2, k=15,
                                   initial_max_displacement=0.65, rng="philox", philox_seed=1)
        eng.particle_displacement(1000)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            eng.particle_displacement(1000)
        b.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / 5
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("FS_SWEEP_CHILD"):
        child()
    else:
        for lpc in ("8", "16", "32"):
            env = dict(os.environ, FS_SWEEP_LPC=lpc, FS_SWEEP_CHILD="1")
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True, text=True, env=env)
            print("LPC", lpc, r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-800:])
