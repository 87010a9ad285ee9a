"""fs_energy_total timing per group size G and tile limit (tuning run for energy.cu's heuristics)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import bench
    import flowstate_b200.MCMC as MC
    out = {}
    for n, B in ((64, 65536), (128, 32768), (256, 16384), (512, 8192), (4096, 1024)):
        pos, L = MC.jittered_lattice(n, 0.5, seed=7, batch=64)
        pos = torch.from_numpy(pos).cuda().repeat(B // 64, 1, 1).contiguous()
        try:
            eng = MC.BatchedMonteCarlo(pos[:64], MC.SimulationBox(L), 1.0, n, rng="philox", **bench.POT)
            eng.total_energy_virial(pos)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                eng.total_energy_virial(pos)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            out[n] = round(bench.energy_flops(B, n) / (ms * 1e-3) / 1e12, 2)
        except Exception as ex:
            out[n] = repr(ex)[:40]
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("FS_TUNE_CHILD"):
        child()
    else:
        for g in ("0", "32", "64"):
            for v2 in ("114688",):
                env = dict(os.environ, FS_ENERGY_G=g, FS_ENERGY_V2MAX=v2, FS_TUNE_CHILD="1")
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True, text=True, env=env)
                print("G", g, "v2max", v2, r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-300:])
