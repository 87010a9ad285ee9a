"""Per-kernel device time of one flow pass (torch.profiler / CUPTI), for a bench workload.
usage: kernel_times.py [workload] [rows] [tf32|fp32] [density|sampling]"""
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench
import flowstate_b200.normflows as NF

name = sys.argv[1] if len(sys.argv) > 1 else "alg1_n32"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
w = dict(bench.WORKLOADS[name])
bound = float(np.float32(np.sqrt(w["n"] / w["rho"]))) / 2
m = bench.build_flow(NF, w, bound, "cuda").cuda().eval()
m.precision = sys.argv[3] if len(sys.argv) > 3 else "tf32"
x = (torch.rand(rows, 2 * w["n"], device="cuda") * 2 - 1) * bound
direction = sys.argv[4] if len(sys.argv) > 4 else "density"      # or "sampling"
z = m.q0(rows)
run = (lambda: m.log_prob(x)) if direction == "density" else (lambda: m.forward(z))
for _ in range(2):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
