"""GPU parity: spline-coupling flow kernels against reference-generated golden vectors
and the oracle.  Tolerance: log-probabilities within 1e-4 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flow_ref as fr

TAGS = ["n3_k3", "n4_k4", "n32_k2", "n4_k23", "n8_h128", "n6_h256"]   # the last two reach the tensor path


def _sd(g, prefix="sd__"):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _build(n, K, blocks, H, nb, bound, device="cuda"):
    import flowstate_b200.normflows as NF
    base = NF.Energy.UniformParticle(n, 2, bound, device=device)
    layers = [NF.flows.CircularCoupledRationalQuadraticSpline(2 * n, blocks, H, range(2 * n), num_bins=nb,
                                                              tail_bound=bound) for _ in range(K)]
    return NF.NormalizingFlow(base, layers)


def _load(g):
    m = _build(int(g["n"]), int(g["K"]), int(g["blocks"]), int(g["H"]), int(g["nb"]), float(g["bound"]))
    m.load_state_dict(_sd(g))
    return m.cuda().eval()


def _precisions(model):
    out = ["fp32"]
    try:
        model.precision = "tf32"
        model.log_prob(torch.zeros(1, 2 * model.q0.n_particles, device="cuda"))
        out.append("tf32")
    except Exception:
        pass
    model.precision = "fp32"
    return out


@pytest.mark.parametrize("tag", TAGS)
def test_golden_flow(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "flow_%s.npz" % tag))
    model = _load(g)
    x = torch.from_numpy(g["x"]).cuda()
    z0 = torch.from_numpy(g["z0"]).cuda()
    bound = float(g["bound"])
    for prec in _precisions(model):
        model.precision = prec
        tol = 1e-4
        lp = model.log_prob(x).cpu().numpy()
        ref = g["log_prob"]
        assert np.array_equal(np.isinf(lp), np.isinf(ref)), prec
        fin = np.isfinite(ref)
        rel = np.abs(lp[fin] - ref[fin]) / np.abs(ref[fin])
        assert rel.max() < tol, (tag, prec, rel.max())
        z, ld = model.inverse_and_log_det(x)
        np.testing.assert_allclose(z.cpu().numpy(), g["inv_z"], rtol=0, atol=2e-4 * bound)
        np.testing.assert_allclose(ld.cpu().numpy(), g["inv_ld"], rtol=1e-4, atol=2e-3)
        xf, ldf = model.forward_and_log_det(z0)
        np.testing.assert_allclose(xf.cpu().numpy(), g["fwd_x"], rtol=0, atol=2e-4 * bound)
        np.testing.assert_allclose(ldf.cpu().numpy(), g["fwd_ld"], rtol=1e-4, atol=2e-3)
        assert torch.equal(model.forward(z0), xf) and torch.equal(model.inverse(x), z)
        # round trip (the reference's own test property, flows/flow_test.py:40-48)
        ok = (x.abs() <= bound).all(dim=1)
        xr, ldr = model.forward_and_log_det(z)
        assert (xr[ok] - x[ok]).abs().max().item() < 2e-3 * bound
        assert (ldr + ld)[ok].abs().max().item() < 5e-3
    model.precision = "fp32"
    # single layer through the nn.Module interface of the layer itself (FP32 conditioner: per-layer tolerances)
    model.flows[int(g["K"]) - 1].precision = "fp32"
    y, ldl = model.flows[int(g["K"]) - 1].inverse(x)
    np.testing.assert_allclose(y.cpu().numpy(), g["lastlayer_inv"], rtol=0, atol=2e-5 * bound)
    np.testing.assert_allclose(ldl.cpu().numpy(), g["lastlayer_ld"], rtol=1e-4, atol=1e-4)
    yb, ldb = model.flows[int(g["K"]) - 1].forward(y)
    assert (yb[ok] - x[ok]).abs().max().item() < 1e-3 * bound


def test_against_oracle_alg_shapes():
    """Wider random flows (Alg-1 style bins, Alg-2 style width) against the float64 oracle,
    with the FP32-vs-FP64 self error of the reference arithmetic reported beside it."""
    torch.manual_seed(0)
    # (32, ...): Alg-1 bins; (64, ...): BASELINE config 4 shape (Alg 2: H=128, 2 blocks, 15 bins, N=64);
    # (256, ...): BASELINE config 3 particle count (2N = 512 features > H: multi-piece GEMM0, 8 coordinate chunks);
    # (20, ...), (9, ...): fused spline epilogue with fewer than 32 bins (pad columns) and odd N;
    # (5, ...): odd N, H outside the tensor path; (150, ...): ragged last chunk, warps of a block with unequal
    # chunk counts, 4-byte staging path (N % 4 != 0); (148, ...): ragged chunk on the 16-byte staging path
    for (n, K, blocks, H, nb, sigma, B) in [(32, 3, 4, 256, 32, 0.02, 200), (64, 4, 2, 128, 15, 0.05, 130),
                                            (256, 2, 3, 256, 32, 0.02, 140), (5, 4, 2, 64, 8, 0.05, 33),
                                            (150, 2, 1, 128, 7, 0.03, 37), (148, 2, 1, 64, 6, 0.03, 19),
                                            (20, 2, 1, 256, 12, 0.03, 150), (9, 2, 2, 256, 31, 0.03, 70)]:
        bound = float(np.float32(np.sqrt(n / 0.03))) / 2
        model = _build(n, K, blocks, H, nb, bound, device="cuda")
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(sigma * torch.randn(p.shape, generator=g))
            for name, buf in model.named_buffers():
                if name.endswith("running_mean"):
                    buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
                elif name.endswith("running_var"):
                    buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
        model = model.cuda().eval()
        sd = {k: v.cpu() for k, v in model.state_dict().items()}
        spec = fr.FlowSpec(sd, bound)
        x = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
        with torch.no_grad():
            truth = fr.log_prob(sd, spec, x.double(), dtype=torch.float64).numpy()
            ref32 = fr.log_prob(sd, spec, x, dtype=torch.float32).numpy()
        self_err = np.max(np.abs(ref32 - truth) / np.abs(truth))
        for prec in _precisions(model):
            model.precision = prec
            got = model.log_prob(x.cuda()).cpu().numpy()
            err = np.max(np.abs(got - truth) / np.abs(truth))
            print("N=%d K=%d H=%d %s: kernel err %.2e, reference-fp32 self err %.2e" % (n, K, H, prec, err, self_err))
            assert err < 1e-4, (n, prec, err, self_err)
        z = model.q0(B)
        with torch.no_grad():
            xo, ldo = fr.forward_and_log_det(sd, spec, z.cpu().double(), dtype=torch.float64)
        for prec in _precisions(model):                 # sampling direction, both conditioner paths
            model.precision = prec
            xs, lds = model.forward_and_log_det(z)
            serr = (xs.cpu().double() - xo).abs().max().item() / bound
            print("N=%d K=%d H=%d %s: sample err %.2e of the bound" % (n, K, H, prec, serr))
            # measured on B200: <= 2e-6 (fp32), <= 3e-5 (tensor path) of the bound on these flows
            tol = 2e-5 if prec == "fp32" else 1e-4
            assert serr < tol, (n, prec, serr)
            np.testing.assert_allclose(lds.cpu().numpy(), ldo.numpy(), rtol=1e-3, atol=2e-3 * n)
        model.precision = "fp32"


def test_fused_spline_epilogue_matches_unfused(monkeypatch):
    """H = 256: the tensor path applies the spline in the conditioner's epilogue (fs_flow_coupling);
    FS_NO_FUSE=1 keeps theta + the separate spline kernel.  Both directions must agree to rounding, and the
    single-layer entry point must reproduce the unfused layer."""
    torch.manual_seed(5)
    n, K, blocks, H, nb = 40, 3, 2, 256, 32              # 2 coordinate chunks, ragged (40 = 32 + 8)
    bound = float(np.float32(np.sqrt(n / 0.03))) / 2
    model = _build(n, K, blocks, H, nb, bound, device="cuda")
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.03 * torch.randn(p.shape, generator=g))
    model = model.cuda().eval()
    if "tf32" not in _precisions(model):
        pytest.skip("tensor path unavailable")
    model.precision = "tf32"
    B = 300                                              # 3 row tiles, last one partial
    x = ((torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    x[0, 3] = bound * 1.5                                # outside the spline interval: identity, zero log-det
    z = model.q0(B)
    lq_f = model.log_prob(x)
    xs_f, ld_f = model.forward_and_log_det(z)
    monkeypatch.setenv("FS_NO_FUSE", "1")
    lq_u = model.log_prob(x)
    xs_u, ld_u = model.forward_and_log_det(z)
    monkeypatch.delenv("FS_NO_FUSE")
    np.testing.assert_allclose(lq_f.cpu().numpy(), lq_u.cpu().numpy(), rtol=2e-6, atol=2e-4)
    np.testing.assert_allclose(ld_f.cpu().numpy(), ld_u.cpu().numpy(), rtol=2e-6, atol=2e-4)
    assert (xs_f - xs_u).abs().max().item() < 2e-5 * bound
    # single-layer entry point against theta + oracle-checked spline of the same layer
    pack = model._cuda_pack()
    idf = model.flows[1].prqct.identity_features.tolist()
    trf = model.flows[1].prqct.transform_features.tolist()
    xi = x[:, idf]
    feats = torch.cat([torch.cos(xi * (np.pi / bound)), torch.sin(xi * (np.pi / bound))], 1).contiguous()
    xo, ld = pack.coupling(1, "density", feats, x)
    monkeypatch.setenv("FS_NO_FUSE", "1")
    y_ref, ld_ref = model.flows[1].inverse(x)
    monkeypatch.delenv("FS_NO_FUSE")
    h = n                                                # roll by D/2
    cols = [(t + h) % (2 * n) for t in trf]
    assert (xo[:, cols] - y_ref[:, cols]).abs().max().item() < 2e-5 * bound
    # the row-tiled feature layout of the full passes (FS_FEATURES_TILED) is the same arithmetic: bit-identical
    xo_t, ld_t = pack.coupling(1, "density", pack.tile_features(feats), x, tiled=True)
    assert torch.equal(xo_t, xo) and torch.equal(ld_t, ld)
    xs_r, _ = pack.coupling(1, "sampling", feats, x)
    xs_t, _ = pack.coupling(1, "sampling", pack.tile_features(feats), x, tiled=True)
    assert torch.equal(xs_t, xs_r)


@pytest.mark.parametrize("n,H,nb", [(40, 256, 32), (3, 32, 8), (64, 128, 15)])
def test_prep_kernel_with_staged_tables_is_bit_identical(monkeypatch, n, H, nb):
    """prep_v3 (knot tables staged in shared memory, 32 rows per block) against prep_*_v2 (FS_PREP_V2=1: tables
    gathered from global memory, one warp per row): same elements summed in the same order -> identical bits, in both
    directions and on both conditioner paths, ragged coordinate slices (40 = 32 + 8, 3) and ragged row blocks."""
    bound = float(np.float32(np.sqrt(n / 0.03))) / 2
    model = _build(n, 3, 2, H, nb, bound, device="cuda")
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    model = model.cuda().eval()
    B = 77
    x = ((torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    x[1, 0] = bound * 1.2                                # identity coordinate outside the interval
    z = model.q0(B)
    for prec in _precisions(model):
        model.precision = prec
        monkeypatch.setenv("FS_PREP_V3", "1")
        lq3 = model.log_prob(x)
        xs3, ld3 = model.forward_and_log_det(z)
        monkeypatch.delenv("FS_PREP_V3")
        monkeypatch.setenv("FS_PREP_V2", "1")
        lq2 = model.log_prob(x)
        xs2, ld2 = model.forward_and_log_det(z)
        monkeypatch.delenv("FS_PREP_V2")
        assert torch.equal(lq3, lq2) and torch.equal(xs3, xs2) and torch.equal(ld3, ld2), prec
    model.precision = "fp32"


@pytest.mark.parametrize("n,K,blocks,H,nb,B", [(40, 5, 2, 256, 32, 300), (64, 4, 2, 128, 15, 700), (6, 7, 1, 128, 8, 129),
                                                 (5, 3, 2, 128, 8, 200), (64, 23, 2, 128, 15, 4096)])
def test_layer_parallel_pass_matches_layer_by_layer(monkeypatch, n, K, blocks, H, nb, B):
    """Even N: the conditioners of all K layers run in ONE launch of K x tiles CTAs, the spline chain ordered by
    per-(tile, quadrant, pair) chunk counters in global memory; FS_NO_LP=1 keeps one launch per layer.
    Same kernel arithmetic on the same inputs: coordinates must be bit-identical, log-dets equal up to the order in which
    the per-layer partial sums are added.  N = 5 (odd: the roll swaps the coordinate parity) must fall back silently."""
    bound = float(np.float32(np.sqrt(n / 0.03))) / 2
    model = _build(n, K, blocks, H, nb, bound, device="cuda")
    g = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    model = model.cuda().eval()
    if "tf32" not in _precisions(model):
        pytest.skip("tensor path unavailable")
    model.precision = "tf32"
    x = ((torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    x[2, 1] = bound * 1.3                                # transformed coordinate outside the interval
    z = model.q0(B)
    for _ in range(2):                                   # twice: the chunk counters are re-armed per pass
        z_lp, ld_lp = model.inverse_and_log_det(x)
        xs_lp, lds_lp = model.forward_and_log_det(z)
    # the scheduling hint: "never" = one launch per layer; "prefer" = layer-parallel whenever two whole steps are resident
    lp_auto = model._cuda_pack().uses_layer_parallel(B)
    model.layer_parallel = "never"
    assert not model._cuda_pack().uses_layer_parallel(B)
    z_nv, ld_nv = model.inverse_and_log_det(x)
    model.layer_parallel = "prefer"
    lp_prefer = model._cuda_pack().uses_layer_parallel(B)
    assert lp_prefer or not lp_auto
    if n % 2 == 0:
        assert lp_prefer                                 # every case here keeps two whole steps resident
    z_pf, ld_pf = model.inverse_and_log_det(x)
    model.layer_parallel = "auto"
    assert torch.equal(z_nv, z_lp) and torch.equal(z_pf, z_lp)
    assert (ld_nv - ld_lp).abs().max().item() < 2e-6 * (1.0 + ld_lp.abs().max().item())
    assert (ld_pf - ld_lp).abs().max().item() < 2e-6 * (1.0 + ld_lp.abs().max().item())
    monkeypatch.setenv("FS_NO_LP", "1")
    z_seq, ld_seq = model.inverse_and_log_det(x)
    xs_seq, lds_seq = model.forward_and_log_det(z)
    monkeypatch.delenv("FS_NO_LP")
    assert torch.equal(z_lp, z_seq) and torch.equal(xs_lp, xs_seq)
    scale = 1.0 + ld_seq.abs().max().item()
    assert (ld_lp - ld_seq).abs().max().item() < 2e-6 * scale
    assert (lds_lp - lds_seq).abs().max().item() < 2e-6 * (1.0 + lds_seq.abs().max().item())


def test_sample_and_base_distribution():
    g = torch.Generator().manual_seed(0)
    model = _build(4, 2, 2, 32, 8, 5.0, device="cuda").cuda().eval()
    torch.manual_seed(3)
    s = model.sample(1000)
    assert s.shape == (1000, 8) and s.is_cuda and s.dtype == torch.float32
    # identity-initialised flow = identity map with log-det 0 (wrapper.py:181-185)
    torch.manual_seed(3)
    assert torch.allclose(s, model.q0(1000), atol=1e-5)
    lp = model.log_prob(s)
    assert torch.allclose(lp, torch.full_like(lp, -8 * np.log(10.0)), atol=1e-4)
    out = s.clone()
    out[0, 0] = 5.5
    assert model.log_prob(out)[0].item() == -float("inf")
    del g


def test_pack_follows_parameter_updates():
    model = _build(3, 2, 2, 16, 8, 5.0, device="cuda").cuda().eval()
    x = (torch.rand(8, 6, device="cuda") * 2 - 1) * 5
    a = model.log_prob(x)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    b = model.log_prob(x)
    assert not torch.allclose(a, b)
    model.train()
    with torch.no_grad():
        ref = model.log_prob(x[:8])           # autograd path, BN batch statistics
    assert ref.shape == (8,)
    model.eval()
    c = model.log_prob(x)
    assert torch.isfinite(c).all()
    with pytest.raises(ValueError):
        model.log_prob(torch.zeros(4, 5, device="cuda"))


def test_empty_and_single_row_batches():
    """Empty input (the reference's torch modules return empty tensors) and a single row (one partial tile)."""
    torch.manual_seed(2)
    n, bound = 16, 6.0
    model = _build(n, 2, 1, 256, 32, bound, device="cuda").cuda().eval()
    for prec in _precisions(model):
        model.precision = prec
        e = model.log_prob(torch.empty(0, 2 * n, device="cuda"))
        assert e.shape == (0,)
        z, ld = model.forward_and_log_det(torch.empty(0, 2 * n, device="cuda"))
        assert z.shape == (0, 2 * n) and ld.shape == (0,)
        x1 = (torch.rand(1, 2 * n, device="cuda") * 2 - 1) * bound
        x5 = torch.cat([x1, (torch.rand(4, 2 * n, device="cuda") * 2 - 1) * bound])
        a, b = model.log_prob(x1), model.log_prob(x5)
        assert a.shape == (1,) and torch.isfinite(a).all()
        assert abs(a[0].item() - b[0].item()) <= 1e-6 * abs(b[0].item())      # rows are independent


def test_tensor_path_is_deterministic_run_to_run():
    """The fused kernel has no atomics on its data path and sums the per-row log-det in a fixed order: repeated
    launches on the same input must agree bit for bit (a shared-memory / TMEM hand-off race would not)."""
    torch.manual_seed(4)
    for (n, H, nb, B) in ((32, 256, 32, 1000), (64, 128, 15, 700)):
        bound = float(np.float32(np.sqrt(n / 0.03))) / 2
        model = _build(n, 3, 2, H, nb, bound, device="cuda")
        g = torch.Generator().manual_seed(8)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.03 * torch.randn(p.shape, generator=g))
        model = model.cuda().eval()
        if "tf32" not in _precisions(model):
            pytest.skip("tensor path unavailable")
        model.precision = "tf32"
        x = ((torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound).cuda()
        z = model.q0(B)
        lq0 = model.log_prob(x)
        xs0, ld0 = model.forward_and_log_det(z)
        for _ in range(6):
            assert torch.equal(model.log_prob(x), lq0)
            xs, ld = model.forward_and_log_det(z)
            assert torch.equal(xs, xs0) and torch.equal(ld, ld0)


def test_concurrent_passes_on_two_streams_match_sequential():
    """bench.py samples the next round's proposals on a side stream while the log-density pass runs: both passes share
    the packed weights but use per-stream workspaces; results must equal the sequential ones bit for bit."""
    torch.manual_seed(6)
    n, bound = 32, float(np.float32(np.sqrt(32 / 0.03))) / 2
    model = _build(n, 4, 3, 256, 32, bound, device="cuda")
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.03 * torch.randn(p.shape, generator=g))
    model = model.cuda().eval()
    if "tf32" not in _precisions(model):
        pytest.skip("tensor path unavailable")
    model.precision = "tf32"
    x = ((torch.rand(2048, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    z = model.q0(1024)
    lq_ref = model.log_prob(x)
    xs_ref = model.forward(z)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for _ in range(4):
        with torch.cuda.stream(side):
            xs = model.forward(z)
        lq = model.log_prob(x)
        torch.cuda.synchronize()
        assert torch.equal(lq, lq_ref) and torch.equal(xs, xs_ref)


def test_round_trip_on_the_tensor_path():
    """forward then inverse returns the base sample and cancels the log-determinant (the reference's own test style,
    flows/neural_spline/coupling_test.py:40-59), here through both directions of the fused kernel: the density pass
    sees the same conditioner outputs as the sampling pass that generated the point."""
    torch.manual_seed(7)
    for (n, H, nb) in ((32, 256, 32), (64, 128, 15)):
        bound = float(np.float32(np.sqrt(n / 0.03))) / 2
        model = _build(n, 5, 3, H, nb, bound, device="cuda")
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.04 * torch.randn(p.shape, generator=g))
        model = model.cuda().eval()
        if "tf32" not in _precisions(model):
            pytest.skip("tensor path unavailable")
        model.precision = "tf32"
        z = model.q0(900)
        x, ld_f = model.forward_and_log_det(z)
        z2, ld_i = model.inverse_and_log_det(x)
        inside = (x.abs() <= bound).all(dim=1)
        assert inside.float().mean().item() > 0.99
        assert (z2[inside] - z[inside]).abs().max().item() < 2e-4 * bound
        np.testing.assert_allclose(ld_i[inside].cpu().numpy(), -ld_f[inside].cpu().numpy(), rtol=2e-4, atol=2e-3)


def _perturbed(n, K, blocks, H, nb, sigma, seed=1, bn_scale=1.0, w_scale=None):
    bound = float(np.float32(np.sqrt(n / 0.03))) / 2
    model = _build(n, K, blocks, H, nb, bound, device="cpu")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(sigma * torch.randn(p.shape, generator=g))
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    return model, bound, g


# The three BASELINE flow architectures at their FULL depth (the shapes bench.py times), float64 oracle as truth:
#   configs[1]: N=32,  K=15, 32 blocks, H=256, 32 bins (main_algorithm_1.py:63-70, NUM_BINS passed as num_blocks)
#   configs[2]: N=256, same flow
#   configs[3]: N=64,  K=23, 2 blocks, H=128, 15 bins (main_algorithm_2.py:62-70)
FULL = [("alg1_n32", 32, 15, 32, 256, 32, 0.02, 4096, 512),
        ("alg1_n256", 256, 15, 32, 256, 32, 0.02, 1024, 96),
        ("alg2_n64", 64, 23, 2, 128, 15, 0.05, 4096, 768)]


@pytest.mark.parametrize("tag,n,K,blocks,H,nb,sigma,B,B_oracle", FULL)
def test_tensor_path_full_depth_vs_float64_oracle(tag, n, K, blocks, H, nb, sigma, B, B_oracle):
    """990 (Alg 1) / 138 (Alg 2) dependent FP16-operand GEMMs per pass: log q within 1e-4 relative of the float64
    oracle in both directions, on the batch size of the bench (a subset of rows goes through the CPU oracle; every
    row is checked for finiteness and against the FP32 CUDA-core path)."""
    model, bound, g = _perturbed(n, K, blocks, H, nb, sigma)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    spec = fr.FlowSpec(sd, bound)
    model = model.cuda().eval()
    if "tf32" not in _precisions(model):
        pytest.skip("tensor path unavailable")
    x = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
    z = (torch.rand(B, 2 * n, generator=g) * 2 - 1) * bound
    sel = torch.linspace(0, B - 1, B_oracle).long()
    with torch.no_grad():
        truth = fr.log_prob(sd, spec, x[sel].double(), dtype=torch.float64).numpy()
        ref32 = fr.log_prob(sd, spec, x[sel], dtype=torch.float32).numpy()
        xo, ldo = fr.forward_and_log_det(sd, spec, z[sel].double(), dtype=torch.float64)
    self_err = np.max(np.abs(ref32 - truth) / np.abs(truth))
    out = {}
    for prec in ("fp32", "tf32"):
        model.precision = prec
        lq = model.log_prob(x.cuda())
        xs, lds = model.forward_and_log_det(z.cuda())
        model._cuda_pack().check_nan()
        assert torch.isfinite(lq).all() and torch.isfinite(xs).all() and torch.isfinite(lds).all()
        out[prec] = (lq.cpu().numpy(), xs.cpu().double(), lds.cpu().numpy())
        err = np.max(np.abs(out[prec][0][sel.numpy()] - truth) / np.abs(truth))
        serr = (out[prec][1][sel] - xo).abs().max().item() / bound
        # the sampling pass' log-det enters log q(x) = -D log(2 bound) - log-det: same relative scale as log q
        base = 2 * n * np.log(2 * bound)
        lerr = np.max(np.abs(out[prec][2][sel.numpy()] - ldo.numpy()) / np.abs(base + ldo.numpy()))
        print("%s %s: log q err %.2e (reference-fp32 self err %.2e), sample err %.2e of the bound, sampling log-det err "
              "%.2e" % (tag, prec, err, self_err, serr, lerr))
        assert err < 1e-4, (tag, prec, err, self_err)
        assert lerr < 1e-4, (tag, prec, lerr)
        # measured on B200: fp32 <= 6.6e-6, tensor path 2.4e-5 / 3.6e-5 / 2.7e-4 (alg1_n32 / alg1_n256 / alg2_n64)
        assert serr < (2e-5 if prec == "fp32" else 6e-4), (tag, prec, serr)
    # every row of the batch: tensor path against the FP32 CUDA-core path
    rel = np.abs(out["tf32"][0] - out["fp32"][0]) / np.abs(out["fp32"][0])
    assert rel.max() < 1e-4, (tag, rel.max())


@pytest.mark.parametrize("tag,n,K,blocks,H,nb,sigma,B,B_oracle", FULL)
def test_sampling_pass_log_prob_equals_inverse_pass(tag, n, K, blocks, H, nb, sigma, B, B_oracle):
    """NormalizingFlow.sample_with_log_prob: log q(x) = log q0(z) - log-det of the sampling pass, against what the
    reference's nf_big_move computes for a proposal - log_prob of the sampled x through the inverse pass
    (MCMC/monte_carlo.py:262) - at the full depth of the three BASELINE flows and on the path the bench runs.  The two
    differ by the round-off of inverting the flow; required: within the 1e-4 relative log-density tolerance (and the
    float64 oracle's log_prob of the same x as the arbiter on a subset of rows)."""
    model, bound, g = _perturbed(n, K, blocks, H, nb, sigma)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    spec = fr.FlowSpec(sd, bound)
    model = model.cuda().eval()
    model.q0.device = "cuda"                      # the base distribution samples where it is told to (Energy/Uniform.py:18-22)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    for prec in _precisions(model):
        model.precision = prec
        x, lq_s = model.sample_with_log_prob(B)
        lq_i = model.log_prob(x)
        model._cuda_pack().check_nan()
        assert x.shape == (B, 2 * n) and torch.isfinite(lq_s).all() and torch.isfinite(lq_i).all()
        rel = ((lq_s - lq_i).abs() / lq_i.abs()).max().item()
        sel = torch.linspace(0, B - 1, min(B_oracle, 64)).long()
        with torch.no_grad():
            truth = fr.log_prob(sd, spec, x[sel].cpu().double(), dtype=torch.float64).numpy()
        err = np.max(np.abs(lq_s[sel].cpu().numpy() - truth) / np.abs(truth))
        print("%s %s: sampling-pass log q vs inverse pass %.2e, vs float64 oracle %.2e" % (tag, prec, rel, err))
        assert rel < 1e-4 and err < 1e-4, (tag, prec, rel, err)


def test_fp16_operand_range_is_guarded():
    """FP16 operands overflow at 65504.  A flow whose activations leave that range must not return silently wrong
    numbers from the tensor path: the NaN flag is raised (check_nan -> ValueError, the reference's error for a broken
    spline, utils/splines.py:176-183), while precision="fp32" evaluates the same flow within tolerance.  Tiny weights
    (FP16 subnormal range) must stay inside the tolerance."""
    n, K, blocks, H, nb = 16, 2, 2, 256, 32
    model, bound, g = _perturbed(n, K, blocks, H, nb, 0.02, seed=4)
    x = ((torch.rand(256, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    # (a) huge BatchNorm scale in block 0 of layer 0: relu(s u + o) ~ 1e6 >> 65504
    big = _perturbed(n, K, blocks, H, nb, 0.02, seed=4)[0]
    with torch.no_grad():
        big.flows[0].prqct.transform_net.blocks[0].batch_norm_layers[0].weight.mul_(3e6)
    big = big.cuda().eval()
    if "tf32" not in _precisions(big):
        pytest.skip("tensor path unavailable")
    big.precision = "tf32"
    big.log_prob(x)
    with pytest.raises(ValueError):
        big._cuda_pack().check_nan()
    big.precision = "fp32"
    lq32 = big.log_prob(x)
    big._cuda_pack().check_nan()
    sd = {k: v.cpu() for k, v in big.state_dict().items()}
    with torch.no_grad():
        truth = fr.log_prob(sd, fr.FlowSpec(sd, bound), x[:32].cpu().double(), dtype=torch.float64).numpy()
    assert np.max(np.abs(lq32[:32].cpu().numpy() - truth) / np.abs(truth)) < 1e-4
    # (b) second-linear weights of every block scaled into the FP16 subnormal range (|w| ~ 2e-5 < 6.1e-5): the
    #     operands lose relative precision but the absolute error stays far below the tolerance
    tiny = _perturbed(n, K, blocks, H, nb, 0.02, seed=4)[0]
    with torch.no_grad():
        for f in tiny.flows:
            for blk in f.prqct.transform_net.blocks:
                blk.linear_layers[1].weight.mul_(1e-3)
    tiny = tiny.cuda().eval()
    tiny.precision = "tf32"
    lq = tiny.log_prob(x)
    tiny._cuda_pack().check_nan()
    tiny.precision = "fp32"
    ref = tiny.log_prob(x)
    err = ((lq - ref).abs() / ref.abs()).max().item()
    print("FP16-subnormal weights: tensor path vs fp32 path %.2e" % err)
    assert err < 1e-4


@pytest.mark.parametrize("n,K,blocks,H,nb", [(32, 3, 4, 256, 32), (64, 4, 2, 128, 15), (5, 3, 2, 64, 8), (40, 2, 1, 256, 20)])
def test_device_side_update_equals_a_fresh_pack(n, K, blocks, H, nb):
    """fs_flow_update (the per-cycle refresh of Algorithm 2: optimizer step -> eval) must give what fs_flow_create
    gives on the same parameters, on both conditioner paths and in both directions, without rebuilding the pack."""
    model, bound, g = _perturbed(n, K, blocks, H, nb, 0.03, seed=11)
    model = model.cuda().eval()
    x = ((torch.rand(300, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    z = ((torch.rand(300, 2 * n, generator=g) * 2 - 1) * bound).cuda()
    model.log_prob(x)                                    # pack created on the host from the initial parameters
    pack = model._cuda_pack()
    with torch.no_grad():                                # an "optimizer step": every tensor changes in place
        for p in model.parameters():
            p.add_(0.02 * torch.randn(p.shape, generator=g).cuda())
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.add_(0.05 * torch.randn(buf.shape, generator=g).cuda())
            elif name.endswith("running_var"):
                buf.mul_(1.1)
    fresh = _build(n, K, blocks, H, nb, bound, device="cpu")
    fresh.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    fresh = fresh.cuda().eval()
    for prec in _precisions(fresh):
        model.precision = fresh.precision = prec
        a, b = model.log_prob(x), fresh.log_prob(x)
        assert model._cuda_pack() is pack and getattr(pack, "updates", 0) >= 1      # refreshed in place, not rebuilt
        xa, la = model.forward_and_log_det(z)
        xb, lb = fresh.forward_and_log_det(z)
        err = ((a - b).abs() / b.abs()).max().item()
        print("N=%d H=%d %s: updated vs fresh pack: log q %.2e, sample %.2e" % (n, H, prec, err, (xa - xb).abs().max().item() / bound))
        assert err < 2e-6 and (xa - xb).abs().max().item() < 2e-6 * bound and (la - lb).abs().max().item() < 2e-4
    # train() / eval() round trip keeps the pack (and its device buffers)
    model.train()
    model.eval()
    assert model._cuda_pack() is pack
