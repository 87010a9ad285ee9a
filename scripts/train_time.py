"""Times one forward-KL training step of the Alg-2 flow (K=23, H=128, 2 blocks, 15 bins, N=64, batch 256):
eager autograd vs the CUDA-graph replay of drivers.training.FlowTrainer."""
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
for graph in (False, True):
    model = bench.build_flow(NF, w, bound, "cuda").cuda()
    tr = FlowTrainer(model, 5e-4, 1e-4, 1.0, 256, use_graph=graph)
    x = ((torch.rand(256, 2 * w["n"]) * 2 - 1) * bound).cuda()
    model.train()
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        tr.step(x)
    torch.cuda.synchronize()
    print("graph=%s: %.2f ms per training step" % (graph, (time.perf_counter() - t0) * 100))
