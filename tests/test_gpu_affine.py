"""GPU parity of the affine (RealNVP) coupling layers and periodic flows (SURVEY.md 8 row f3:
NF/normflows/flows/affine/coupling.py:99-268, flows/periodic.py:6-73) against outputs of the reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _stack(g):
    import flowstate_b200.normflows as NF
    D, bound = int(g["D"]), float(g["bound"])
    b = torch.tensor([1.0 if i % 2 == 0 else 0.0 for i in range(D)])
    flows = []
    for i in range(4):
        s_net = NF.nets.MLP([D, 2 * D, D], init_zeros=True)
        t_net = NF.nets.MLP([D, 2 * D, D], init_zeros=True)
        flows.append(NF.flows.MaskedAffineFlow(b if i % 2 == 0 else 1 - b, t_net, s_net))
        flows.append(NF.flows.PeriodicShift(list(range(0, D, 3)), bound=bound, shift=0.37 * (i + 1)))
    model = NF.NormalizingFlow(NF.Energy.UniformParticle(D // 2, 2, bound, device="cuda"), flows)
    model.load_state_dict(_sd(g, "stack_sd__"))          # the reference's state_dict loads unchanged
    return model.cuda().eval()


def test_masked_affine_stack_matches_reference(golden_dir):
    from flowstate_b200 import _lib
    g = np.load(os.path.join(golden_dir, "affine.npz"))
    model = _stack(g)
    l0 = _lib.lib().fs_launch_count()
    with torch.no_grad():
        z = torch.from_numpy(g["stack_z"]).cuda()
        for i, f in enumerate(model.flows):
            z, ld = f(z)
            np.testing.assert_allclose(z.cpu().numpy(), g["stack_fwd_%d" % i], rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(ld.cpu().numpy(), g["stack_fwd_ld"][i], rtol=2e-5, atol=2e-6)
        x = torch.from_numpy(g["stack_x"]).cuda()
        for j, i in enumerate(range(len(model.flows) - 1, -1, -1)):
            x, ld = model.flows[i].inverse(x)
            np.testing.assert_allclose(x.cpu().numpy(), g["stack_inv_%d" % i], rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(ld.cpu().numpy(), g["stack_inv_ld"][j], rtol=2e-5, atol=2e-6)
        lp = model.log_prob(torch.from_numpy(g["stack_x"]).cuda())
        ref = g["stack_log_prob"]
        assert np.array_equal(np.isinf(lp.cpu().numpy()), np.isinf(ref))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(lp.cpu().numpy()[fin], ref[fin], rtol=1e-4, atol=1e-5)
    assert _lib.lib().fs_launch_count() - l0 >= 3 * len(model.flows)          # the CUDA kernels ran, not the torch path
    # non-finite parameters -> NaN like the reference (coupling.py:199-202)
    with torch.no_grad():
        model.flows[0].s.net[-1].weight.copy_(torch.from_numpy(g["nan_last_w"]).cuda())
        y, ld = model.flows[0](torch.from_numpy(g["nan_in"]).cuda())
    assert np.array_equal(np.isnan(y.cpu().numpy()), np.isnan(g["nan_out"]))
    assert np.array_equal(np.isnan(ld.cpu().numpy()), np.isnan(g["nan_ld"]))
    # with autograd the torch expressions run and agree with the kernel path
    zg = torch.from_numpy(g["stack_z"]).cuda().requires_grad_(True)
    yg, lg = model.flows[2](zg)
    (yg.sum() + lg.sum()).backward()
    assert torch.isfinite(zg.grad).all()
    with torch.no_grad():
        yk, lk = model.flows[2](zg.detach())
    assert (yg.detach() - yk).abs().max().item() < 1e-5 and (lg.detach() - lk).abs().max().item() < 1e-5


def test_affine_coupling_blocks_and_wrap_match_reference(golden_dir):
    import flowstate_b200.normflows as NF
    g = np.load(os.path.join(golden_dir, "affine.npz"))
    D, bound = int(g["D"]), float(g["bound"])
    z = torch.from_numpy(g["stack_z"]).cuda()
    x = torch.from_numpy(g["stack_x"]).cuda()
    with torch.no_grad():
        for sm in ("exp", "sigmoid", "sigmoid_inv"):
            for mode in ("channel", "channel_inv"):
                tag = "blk_%s_%s" % (sm, mode)
                blk = NF.flows.AffineCouplingBlock(NF.nets.MLP([D // 2, 16, D], init_zeros=True), scale=True,
                                                   scale_map=sm, split_mode=mode)
                blk.load_state_dict(_sd(g, tag + "_sd__"))
                blk = blk.cuda().eval()
                yf, lf = blk(z)
                yi, li = blk.inverse(x)
                np.testing.assert_allclose(yf.cpu().numpy(), g[tag + "_fwd"], rtol=2e-5, atol=2e-5)
                np.testing.assert_allclose(lf.cpu().numpy(), g[tag + "_fwd_ld"], rtol=2e-5, atol=2e-5)
                np.testing.assert_allclose(yi.cpu().numpy(), g[tag + "_inv"], rtol=2e-5, atol=2e-5)
                np.testing.assert_allclose(li.cpu().numpy(), g[tag + "_inv_ld"], rtol=2e-5, atol=2e-5)
                back, lb = blk.inverse(yf)                         # round trip (the reference's own test property)
                assert (back - z).abs().max().item() < 2e-4 and (lb + lf).abs().max().item() < 2e-4
        blk = NF.flows.AffineCouplingBlock(NF.nets.MLP([D // 2, 16, D // 2], init_zeros=True), scale=False)
        blk.load_state_dict(_sd(g, "blk_noscale_sd__"))
        blk = blk.cuda().eval()
        yf, lf = blk(z)
        np.testing.assert_allclose(yf.cpu().numpy(), g["blk_noscale_fwd"], rtol=2e-5, atol=2e-5)
        assert float(lf.abs().max()) == 0.0
        wrap = NF.flows.PeriodicWrap(list(range(1, D, 2)), bound=bound)
        out, ld = wrap.inverse(torch.from_numpy(g["wrap_in"]).cuda())
        np.testing.assert_allclose(out.cpu().numpy(), g["wrap_inv"], rtol=1e-6, atol=2e-6)
        assert torch.equal(wrap.forward(z)[0], z) and float(ld.abs().max()) == 0.0
