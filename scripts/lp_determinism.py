"""development: run-to-run determinism of the layer-parallel flow passes (run under gpurun); reports where repeats differ"""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_gpu_flow import _build

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfgs = ((32, 256, 32, 1000, 3), (64, 128, 15, 700, 3), (32, 256, 32, 4096, 15), (64, 128, 32, 700, 3), (64, 256, 15, 700, 3))
if len(sys.argv) > 2:
    cfgs = [cfgs[int(a)] for a in sys.argv[2:]]
torch.manual_seed(4)
for (n, H, nb, B, K) in cfgs:
    bound = float(np.float32(np.sqrt(n / 0.03))) / 2
    model = _build(n, K, 2, H, nb, bound, device="cuda")
    g = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.03 * torch.randn(p.shape, generator=g))
    model = model.cuda().eval()
    model.precision = "tf32"
    x = ((torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    z = model.q0(B)
    lq0 = model.log_prob(x)
    xs0, ld0 = model.forward_and_log_det(z)
    bad_lq = bad_xs = 0
    where = []
    for it in range(iters):
        lq = model.log_prob(x)
        xs, ld = model.forward_and_log_det(z)
        if not torch.equal(lq, lq0):
            bad_lq += 1
            rows = torch.nonzero(lq != lq0, as_tuple=True)[0]
            where.append(("log_prob", it, rows[:4].tolist(), int(rows.numel()), (lq - lq0)[rows[:3]].tolist()))
        if not torch.equal(xs, xs0):
            bad_xs += 1
            rows, cols = torch.nonzero(xs != xs0, as_tuple=True)
            where.append(("sample", it, rows[:4].tolist(), sorted(set(cols.tolist())), int(rows.numel()),
                          xs[rows[:3], cols[:3]].tolist(), xs0[rows[:3], cols[:3]].tolist(),
                          "ld differs" if not torch.equal(ld, ld0) else "ld equal"))
    print("n=%d H=%d nb=%d K=%d B=%d: log_prob repeats differing %d / %d, sample repeats differing %d / %d"
          % (n, H, nb, K, B, bad_lq, iters, bad_xs, iters))
    for w in where[:12]:
        print("   ", w)
