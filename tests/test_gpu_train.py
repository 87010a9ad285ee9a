"""GPU: the training path on the device - hand-written spline forward / backward kernels (fs_spline_train_fwd / _bwd)
against autograd through the torch restatement of utils/splines.py, the forward-KL gradient of a whole flow against the
oracle's autograd on the CPU, and the CUDA-graph trainer against eager steps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(n, K, blocks, H, nb, bound):
    import flowstate_b200.normflows as NF
    base = NF.Energy.UniformParticle(n, 2, bound, device="cuda")
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, blocks, H, range(2 * n), num_bins=nb,
                                                              tail_bound=bound) for _ in range(K)]
    return NF.NormalizingFlow(base, layers)


@pytest.mark.parametrize("nb,N,B,shared,std,tol", [(15, 64, 256, False, 0.7, 2e-4), (32, 32, 100, False, 0.7, 2e-4),
                                                    (8, 5, 33, True, 1.0, 2e-4), (15, 64, 256, True, 1.0, 2e-4),
                                                    # stress: logits of std 3 squeeze bins to the 1e-3 minimum width,
                                                    # slopes ~3e3: float32 against float64 is good to ~1e-2 there (torch float32: 2e-3 .. 4e-3)
                                                    (15, 64, 256, False, 3.0, 2e-2), (32, 32, 100, False, 3.0, 2e-2)])
def test_fused_spline_matches_torch_autograd(nb, N, B, shared, std, tol):
    from flowstate_b200.normflows import _spline_torch as st
    g = torch.Generator().manual_seed(3)
    bound, P = 7.5, 3 * nb + 1
    x = ((torch.rand(B, N, generator=g) * 2 - 1) * bound).cuda()
    x[0, 0] = bound                      # on the upper edge: last bin
    x[1, 1] = -bound
    x[2, 2] = 1.3 * bound                # outside: identity, zero log-det, gradient 1
    theta = (torch.randn((N, P) if shared else (B, N, P), generator=g) * std).cuda()
    scale = 1.0   # (the 1 / sqrt(hidden) of the conditional spline is exercised by the forward-KL test below)
    gy = torch.randn(B, N, generator=g).cuda()
    gl = torch.randn(B, N, generator=g).cuda()

    def ref(xx, th):
        t = th[None].expand(B, N, P) if shared else th
        return st.spline(xx, t[..., :nb] * scale, t[..., nb:2 * nb] * scale, t[..., 2 * nb:], bound, False)
    def run(fn, xx, th):
        xx, th = xx.clone().requires_grad_(True), th.clone().requires_grad_(True)
        yy, ll = fn(xx, th)
        gxx, gth = torch.autograd.grad((yy * gy.to(yy.dtype)).sum() + (ll * gl.to(yy.dtype)).sum(), (xx, th))
        return [t.detach().double() for t in (yy, ll, gxx, gth)]
    truth = run(ref, x.double(), theta.double())                        # float64 autograd through the torch restatement
    ref32 = run(ref, x, theta)                                          # the reference's own float32 arithmetic
    fused = run(lambda xx, th: st.fused_spline(xx, th, bound, nb, scale), x, theta)
    # yardstick: the kernels must be as accurate as the reference's float32 arithmetic on the same inputs (steep bins
    # - widths down to 1e-3 of the interval - amplify float32 rounding for any implementation)
    for name, t64, t32, tf in zip(("y", "log-det", "grad_x", "grad_theta"), truth, ref32, fused):
        sc = max(1.0, t64.abs().max().item())
        e32, ef = (t32 - t64).abs().max().item() / sc, (tf - t64).abs().max().item() / sc
        print("nb=%d shared=%s std=%g %-10s: fused %.2e, torch float32 %.2e (of scale %.3g)" % (nb, shared, std, name, ef, e32, sc))
        assert ef <= max(3.0 * e32, tol), (name, ef, e32)


def test_forward_kld_gradient_fused_vs_torch_path():
    """Loss and gradient of NormalizingFlow.forward_kld in train mode (BatchNorm batch statistics) with the fused spline
    kernels, against the same pass through the torch-op splines."""
    torch.manual_seed(0)
    n, K, blocks, H, nb, bound = 8, 3, 2, 32, 8, 8.0
    model = _build(n, K, blocks, H, nb, bound)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    model = model.cuda().train()
    x = ((torch.rand(64, 2 * n) * 2 - 1) * bound).cuda()
    losses, grads = {}, {}
    for fused in (True, False):
        for f in model.flows:
            f.fused_training = fused
        model.zero_grad()
        # same BatchNorm running-stat side effects in both runs do not matter: batch statistics are used
        loss = model.forward_kld(x)
        loss.backward()
        losses[fused] = float(loss.detach())
        grads[fused] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()
    assert abs(losses[True] - losses[False]) < 1e-4 * max(1.0, abs(losses[False]))
    scale = grads[False].abs().max().item()
    err = (grads[True] - grads[False]).abs().max().item()
    print("forward_kld: loss %.6f / %.6f, gradient err %.2e of %.2e" % (losses[True], losses[False], err, scale))
    assert err < 2e-3 * scale
    # (the torch path used as the yardstick is itself pinned on the CPU against the reference, tests/test_host_cpu.py)


def test_graph_trainer_matches_eager_steps():
    from flowstate_b200.drivers.training import FlowTrainer
    n, K, blocks, H, nb, bound = 6, 2, 2, 32, 8, 6.0
    xs = [((torch.rand(48, 2 * n, generator=torch.Generator().manual_seed(10 + i)) * 2 - 1) * bound).cuda()
          for i in range(4)]
    out = {}
    for graph in (False, True):
        torch.manual_seed(1)
        model = _build(n, K, blocks, H, nb, bound)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.05 * torch.randn_like(p))
        model = model.cuda().train()
        tr = FlowTrainer(model, 1e-3, 1e-4, 1.0, 48, use_graph=graph, native=False)
        losses = [tr.step(x) for x in xs]
        out[graph] = (losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone())
    la, lb = out[False][0], out[True][0]
    assert all(abs(a - b) < 1e-4 * max(1.0, abs(a)) for a, b in zip(la, lb)), (la, lb)
    assert (out[False][1] - out[True][1]).abs().max().item() < 1e-4


def _perturbed(n, K, blocks, H, nb, bound, seed, sigma=0.05):
    torch.manual_seed(seed)
    model = _build(n, K, blocks, H, nb, bound)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(sigma * torch.randn_like(p))
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn_like(buf))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand_like(buf))
    return model.cuda().train()


@pytest.mark.parametrize("n,K,blocks,H,nb,B", [(8, 3, 2, 32, 8, 100), (64, 4, 2, 128, 15, 256), (32, 2, 3, 64, 32, 37)])
def test_native_training_step_matches_autograd(n, K, blocks, H, nb, B):
    """fs_train_forward_kld (the whole forward-KL step in hand-written kernels) against loss.backward() through the
    module tree: loss, the gradient of every parameter, BatchNorm running statistics and counters; and bit-identical
    when repeated.  Tolerance: 2e-4 of the largest gradient entry of a tensor's group (float32, different summation
    order; the autograd path it is compared with is pinned on the reference in tests/test_host_cpu.py)."""
    from flowstate_b200.drivers import _train_native as tn
    bound = float(np.float32(np.sqrt(n / 0.3))) / 2
    model = _perturbed(n, K, blocks, H, nb, bound, seed=5)
    assert tn.supported(model)
    x = ((torch.rand(B, 2 * n, generator=torch.Generator().manual_seed(2)) * 2 - 1) * bound).cuda()
    x[0, 0] = 1.2 * bound                 # a coordinate outside the box: identity, zero log-det
    bufs0 = {k: v.clone() for k, v in model.named_buffers()}
    model.zero_grad()
    loss_ref = model.forward_kld(x)
    loss_ref.backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    bufs_ref = {k: v.clone() for k, v in model.named_buffers()}
    with torch.no_grad():
        for k, v in model.named_buffers():
            v.copy_(bufs0[k])
    for p in model.parameters():
        if p.grad is not None:
            p.grad = torch.full_like(p.grad, 7.0)          # the engine writes, it does not accumulate
    eng = tn.NativeForwardKL(model)
    loss = eng.step(x).clone()
    got = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    loss_ref = loss_ref.detach()
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * max(1.0, abs(float(loss_ref))), (float(loss), float(loss_ref))
    assert set(got) == set(ref)
    worst = 0.0
    top = max(v.abs().max().item() for v in ref.values())
    for k in ref:
        # (a Linear bias in front of a BatchNorm has a gradient of exactly zero up to rounding: judged on the absolute
        # difference, like every tensor whose gradient is below 1e-3 of the largest one)
        sc = max(ref[k].abs().max().item(), 1e-3 * top)
        err = (got[k] - ref[k]).abs().max().item() / sc
        worst = max(worst, err)
        assert err < 2e-4, (k, err, sc)
    for k, v in model.named_buffers():
        if v.dtype.is_floating_point:
            assert (v - bufs_ref[k]).abs().max().item() < 1e-5 * max(1.0, bufs_ref[k].abs().max().item()), k
        else:
            assert torch.equal(v, bufs_ref[k]), k
    print("native step n=%d K=%d H=%d: loss %.6f vs %.6f, worst relative gradient error %.2e" %
          (n, K, H, float(loss), float(loss_ref), worst))
    eng.step(x, update_running=False)
    again = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    assert all(torch.equal(again[k], got[k]) for k in got)


def test_native_trainer_matches_autograd_trainer():
    from flowstate_b200.drivers.training import FlowTrainer
    n, K, blocks, H, nb, bound = 8, 3, 2, 32, 8, 6.0
    xs = [((torch.rand(48, 2 * n, generator=torch.Generator().manual_seed(20 + i)) * 2 - 1) * bound).cuda()
          for i in range(4)]
    out = {}
    for native in (False, True):
        model = _perturbed(n, K, blocks, H, nb, bound, seed=1)
        tr = FlowTrainer(model, 1e-3, 1e-4, 1.0, 48, use_graph=False, native=native)
        losses = [tr.step(x) for x in xs]
        assert (tr.native is not None) == native
        out[native] = (losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone(),
                       torch.cat([b.detach().double().reshape(-1) for b in model.buffers()]).clone())
    la, lb = out[False][0], out[True][0]
    assert all(abs(a - b) < 1e-4 * max(1.0, abs(a)) for a, b in zip(la, lb)), (la, lb)
    assert (out[False][1] - out[True][1]).abs().max().item() < 2e-4
    assert (out[False][2] - out[True][2]).abs().max().item() < 1e-4
    # eval-mode sampling after the native steps sees the trained weights (device re-pack)
    model.eval()
    assert torch.isfinite(model.log_prob(xs[0])).all()


def test_flat_adam_matches_torch_adam():
    """fs_adam_step against torch.optim.Adam (the optimizer of both drivers, main_algorithm_2.py:440): several steps
    with L2 weight decay on a ragged-length vector, a step skipped by the device-side flag and one skipped by a
    non-finite loss (neither moves the parameters, the moments or the step count), then a fresh optimizer."""
    import flowstate_b200._lib as lib
    n = 100003
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * (0.1 + i) for i in range(6)]
    lr, wd = 5.435e-4, 9.586e-5
    ref = torch.nn.Parameter(p0.clone().cuda())
    opt = torch.optim.Adam([ref], lr=lr, weight_decay=wd)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    st = torch.zeros(4, device="cuda")
    one, nan = torch.ones(1, device="cuda"), torch.full((1,), float("nan"), device="cuda")
    fin = torch.zeros(1, device="cuda")

    def step(gr, skip=None, loss=None):
        lib.check(lib.lib().fs_adam_step(lib.ptr(p), lib.ptr(gr), lib.ptr(m), lib.ptr(v), n, lib.ptr(st), lib.ptr(skip),
                                         lib.ptr(loss), lr, 0.9, 0.999, 1e-8, wd, lib.stream_ptr()))
    for i, gr in enumerate(grads):
        gr = gr.cuda()
        if i == 2:
            before = (p.clone(), m.clone(), v.clone(), st.clone())
            step(gr, skip=one)
            step(gr, loss=nan)
            assert torch.equal(p, before[0]) and torch.equal(m, before[1]) and torch.equal(v, before[2])
            assert st[0].item() == before[3][0].item() and st[1].item() == 0
        step(gr, skip=None if i % 2 else torch.zeros(1, device="cuda"), loss=fin)
        ref.grad = gr.clone()
        opt.step()
        assert st[1].item() == 1 and st[0].item() == i + 1
        err = (p - ref.detach()).abs().max().item()
        assert err < 2e-6, (i, err)
    # a fresh optimizer = zeroed state
    m.zero_(); v.zero_(); st.zero_()
    opt = torch.optim.Adam([ref], lr=lr, weight_decay=wd)
    with torch.no_grad():
        ref.copy_(p)
    step(grads[0].cuda())
    ref.grad = grads[0].cuda()
    opt.step()
    assert (p - ref.detach()).abs().max().item() < 2e-6


def test_trainer_flat_adam_matches_torch_optimizer():
    """FlowTrainer with the flat Adam (parameters re-pointed into one buffer) against the same trainer driving
    torch.optim.Adam: losses and parameters after four steps over two optimizer lifetimes; state_dict keys, shapes and
    Parameter objects are untouched; a non-finite batch is skipped; sync=False returns device losses."""
    from flowstate_b200.drivers.training import FlowTrainer
    n, K, blocks, H, nb, bound = 8, 3, 2, 32, 8, 6.0
    xs = [((torch.rand(48, 2 * n, generator=torch.Generator().manual_seed(30 + i)) * 2 - 1) * bound).cuda()
          for i in range(4)]
    out = {}
    for flat in (False, True):
        model = _perturbed(n, K, blocks, H, nb, bound, seed=2)
        keys0 = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        ids0 = [id(p) for p in model.parameters()]
        tr = FlowTrainer(model, 1e-3, 1e-4, 1.0, 48, use_graph=False, native=True, native_adam=flat)
        losses = [tr.step(xs[0]), tr.step(xs[1])]
        tr.fresh_optimizer()
        if flat:
            l2 = tr.step(xs[2], sync=False)
            l3 = tr.step(xs[3], sync=False)
            assert l2.is_cuda and l2.numel() == 1
            losses += [float(l2), float(l3)]
        else:
            losses += [tr.step(xs[2]), tr.step(xs[3])]
        assert (tr.flat_p is not None) == flat
        assert keys0 == [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        assert ids0 == [id(p) for p in model.parameters()]
        out[flat] = (losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone())
        model.eval()
        assert torch.isfinite(model.log_prob(xs[0])).all()      # the inference pack follows the flat update
        lq = model.log_prob(xs[0])
        model.repack()
        assert torch.equal(lq, model.log_prob(xs[0]))
        # a non-finite loss skips the step (main_algorithm_2.py:449); last, because the NaN batch also reaches the
        # BatchNorm running statistics, exactly as it would in the reference's forward pass
        model.train()
        bad = xs[2].clone()
        bad[0, 0] = float("nan")
        before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
        assert tr.step(bad) is None
        assert torch.equal(before, torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    la, lb = out[False][0], out[True][0]
    assert all(abs(a - b) < 1e-5 * max(1.0, abs(a)) for a, b in zip(la, lb)), (la, lb)
    # Adam divides by sqrt(v): an element whose gradient is rounding noise (a Linear bias in front of a BatchNorm has
    # none at all mathematically) takes an O(lr) step in a direction set by the last bits of that gradient, so two
    # runs agree element-wise only to a fraction of lr = 1e-3 (the same bound as the autograd-trainer test above); the
    # update rule itself is pinned to 2e-6 by test_flat_adam_matches_torch_adam
    assert (out[False][1] - out[True][1]).abs().max().item() < 2e-4
