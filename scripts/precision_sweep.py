"""Log-density error of the FP32 and the tensor path against the float64 oracle as the weights move away from the\nidentity initialisation (sigma = std of the perturbation); run from the repo root."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import test_gpu_flow as T
from oracle import flow_ref as fr
for sigma in (0.05, 0.1, 0.2, 0.4):
    for (n, K, blocks, H, nb) in ((32, 3, 8, 256, 32), (64, 4, 2, 128, 15)):
        bound = float(np.float32(np.sqrt(n / 0.03))) / 2
        model = T._build(n, K, blocks, H, nb, bound, device="cuda")
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(sigma * torch.randn(p.shape, generator=g))
            for name, buf in model.named_buffers():
                if name.endswith("running_mean"): buf.copy_(0.3 * torch.randn(buf.shape, generator=g))
                elif name.endswith("running_var"): buf.copy_(0.2 + 2 * torch.rand(buf.shape, generator=g))
        model = model.cuda().eval()
        sd = {k: v.cpu() for k, v in model.state_dict().items()}
        spec = fr.FlowSpec(sd, bound)
        x = (torch.rand(300, 2 * n, generator=g) * 2 - 1) * bound
        with torch.no_grad():
            truth = fr.log_prob(sd, spec, x.double(), dtype=torch.float64).numpy()
            ref32 = fr.log_prob(sd, spec, x, dtype=torch.float32).numpy()
        out = []
        for prec in ("fp32", "tf32"):
            model.precision = prec
            got = model.log_prob(x.cuda()).cpu().numpy()
            out.append(np.max(np.abs(got - truth) / np.abs(truth)))
        print("sigma %.2f N=%d H=%d: fp32 %.2e  tensor %.2e  ref-fp32 %.2e  |logq| %.1f" % (sigma, n, H, out[0], out[1], np.max(np.abs(ref32 - truth) / np.abs(truth)), np.mean(np.abs(truth))))
