"""Times one forward-KL training step of the Alg-2 flow (K=23, H=128, 2 blocks, 15 bins, N=64, batch 256):
eager autograd vs the CUDA-graph replay vs the hand-written step (fs_train_forward_kld) of drivers.training.FlowTrainer;
with --native-only just the latter (for an ncu launch list)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import flowstate_b200.normflows as NF  # noqa: E402
from flowstate_b200.drivers.training import FlowTrainer  # noqa: E402

w = bench.WORKLOADS["alg2_n64"]
bound = float(np.float32(np.sqrt(w["n"] / w["rho"]))) / 2
modes = [("native", True, True)] if "--native-only" in sys.argv else \
    [("eager autograd", False, False), ("CUDA graph", True, False), ("native", True, True)]
for name, graph, native in modes:
    model = bench.build_flow(NF, w, bound, "cuda").cuda()
    tr = FlowTrainer(model, 5e-4, 1e-4, 1.0, 256, use_graph=graph, native=native)
    x = ((torch.rand(256, 2 * w["n"]) * 2 - 1) * bound).cuda()
    model.train()
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        tr.step(x)
    torch.cuda.synchronize()
    print("%s: %.2f ms per training step" % (name, (time.perf_counter() - t0) * 100))
    if native:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tr.native.step(x, update_running=False)
        e1.record()
        torch.cuda.synchronize()
        print("   fs_train_forward_kld alone: %.3f ms (device)" % (e0.elapsed_time(e1) / 10))
