"""Development check of the tensor-core conditioner against the FP32 path (run on a B200)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import flowstate_b200.normflows as NF


def build(n, K, blocks, H, nb, sigma, bound):
    torch.manual_seed(0)
    m = NF.NormalizingFlow(NF.Energy.UniformParticle(n, 2, bound, device="cuda"),
                           [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, blocks, H, range(2 * n), num_bins=nb,
                                                                            tail_bound=bound) for _ in range(K)])
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(sigma * torch.randn(p.shape, generator=g))
        for name, buf in m.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    return m.cuda().eval()


cases = [(8, 1, 1, 128, 8, 0.05, 200), (32, 2, 3, 256, 32, 0.02, 300), (16, 3, 2, 128, 15, 0.05, 1000),
         (160, 1, 2, 256, 8, 0.02, 130)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for (n, K, blocks, H, nb, sigma, B) in cases:
    bound = float(np.sqrt(n / 0.03)) / 2
    m = build(n, K, blocks, H, nb, sigma, bound)
    x = (torch.rand(B, 2 * n, device="cuda") * 2 - 1) * bound
    m.precision = "fp32"
    a = m.log_prob(x)
    torch.cuda.synchronize()
    m.precision = "tf32"
    t0 = time.time()
    b = m.log_prob(x)
    torch.cuda.synchronize()
    rel = ((a - b).abs() / a.abs()).max().item()
    print("N=%d K=%d blocks=%d H=%d nb=%d B=%d: max rel |fp32 - tf32| = %.3e (%.1f ms)  sample %.5f %.5f"
          % (n, K, blocks, H, nb, B, rel, 1e3 * (time.time() - t0), a[0].item(), b[0].item()), flush=True)
    z = m.q0(B)
    m.precision = "fp32"
    xa = m.forward(z)
    m.precision = "tf32"
    xb = m.forward(z)
    print("   forward max abs diff / bound = %.3e" % ((xa - xb).abs().max().item() / bound), flush=True)
print("done")
