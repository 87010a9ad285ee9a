"""Prints where the tensor-core conditioner's roles wait (FS_TC_DEBUG=1).  The in-kernel timers are compiled in only when
the library is built with FS_TC_TIMERS=1:  FS_TC_TIMERS=1 python -c "from flowstate_b200 import build; build.build(force=True)"
(rebuild without it afterwards: the clock reads slow the MMA-issuing warp down)."""
import ctypes as C
import os
import sys

os.environ["FS_TC_DEBUG"] = "1"
import numpy as np
import torch

sys.path.insert(0, ".")
import bench
import flowstate_b200.normflows as NF
from flowstate_b200 import _lib

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "alg1_n32"])
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
w["K"] = 2
bound = float(np.float32(np.sqrt(w["n"] / w["rho"]))) / 2
m = bench.build_flow(NF, w, bound, "cuda").cuda().eval()
m.precision = "tf32"
x = (torch.rand(rows, 2 * w["n"], device="cuda") * 2 - 1) * bound
for _ in range(3):
    m.log_prob(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.log_prob(x)
e1.record()
torch.cuda.synchronize()
print("pass of K=2 layers, %d rows: %.3f ms" % (rows, e0.elapsed_time(e1)))
n = (rows + 127) // 128
buf = (C.c_longlong * (16 * n))()
got = _lib.lib().fs_tc_debug_read(buf, n)
a = np.array(buf[: 16 * got]).reshape(got, 16)
names = ["producer wait-empty", "mma wait-operand", "mma wait-weights", "mma total", "epi wait-accum", "epi total", "mma issue+commit", "epi residual half 0", "epi residual half 1", "epi relu (2 halves)", "epi final chunks", "fused A: loads+signal | T2: features packed", "fused A: bias+softmax | T2: first residual done", "fused A: search+derivs | T2: trunk done", "mma final phase", "mma final: wait drained acc"]
for i, nm in enumerate(names):
    print("%-48s mean %10.0f  min %10.0f  max %10.0f clk" % (nm, a[:, i].mean(), a[:, i].min(), a[:, i].max()))
