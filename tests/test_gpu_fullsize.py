"""GPU tests at BASELINE.json's FULL sizes.  The CPU oracle does not
finish these in seconds, so the checks are size-independent properties of the path:

* Sinkhorn-Knopp output: rows of Q sum to 1, prototype marginals are uniform (N/K per column);
* swapped-prediction gradient: every row of dS sums to zero (softmax minus a distribution), hence
  the bias gradient sums to zero; the loss is finite and the same for the fp16 and bf16x3 score GEMMs;
* linearity of the per-resolution projection: Z of the all-pixel path equals the projection of the
  gathered upsampled+concatenated vectors on sampled pixels (two different kernel routes);
* label maps: labels == arg-max of the returned codes, first index on ties; k-means assignment equals
  an fp64 distance arg-min wherever the top-2 margin is not a rounding tie.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _draws(E, b, d, n_layers, hw, npatch, seed):
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)

    def view():
        return E.ViewDraws(layer_no=[int(rs.randint(n_layers)) for _ in range(b)],
                           pert_z=torch.randn(b, 2 * n_layers, d, generator=g),
                           angle=[float(rs.uniform(-10, 10)) for _ in range(b)],
                           flip=[bool(rs.rand() < 0.5) for _ in range(b)])
    return E.StepDraws(z=torch.randn(b, d, generator=g), view_s=view(), view_t=view(),
                       perms=[[torch.randperm(hw, generator=g) for _ in range(b)] for _ in range(npatch)])


def _generator(size, seed=42, channels=None):
    from ganecdotes_b200.stylegan2.model import Generator
    torch.manual_seed(seed)
    g = Generator(size, 512, 8) if channels is None else Generator(size, 512, 8, channels=channels)
    return g.cuda()


def _head(E, hlen, c, k, seed=1, **kw):
    torch.manual_seed(seed)
    proj = torch.nn.Linear(hlen, c, bias=False)
    proto = torch.nn.Linear(c, k)
    return E.SwavHead(proj.weight.data.cuda(), proto.weight.data.cuda(), proto.bias.data.cuda(), 0.01, 0.9, 0.01, 3, 1,
                      **kw)


@pytest.fixture(scope="module")
def ffhq():
    from ganecdotes_b200.hfc_with_swav import engine as E
    gen = _generator(256)
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(1024, 512, generator=torch.Generator().manual_seed(3)).cuda()).mean(0, keepdim=True)
    return E, gen, mean_latent


def test_ffhq256_full_step_properties(ffhq):
    """config 2 shape per GPU, 2 latents: D=5376, C=512, K=5000, 5 patches x 20000 px, at the SHIPPED
    hyper-parameters of hfc_with_swav_ffhq_config.py (eps 0.005, temperature 0.01: T/eps = 2 -> swav_loss_pow_kernel)"""
    E, gen, mean_latent = ffhq
    cfg = E.StepConfig(hlen=5376, patch_size=20000, num_patches=5, niters=10, eps=0.005, temperature=0.01,
                       truncation=0.7, perturb_std=[1.0] * 6)
    draws = _draws(E, 2, 512, 6, 65536, 5, 7)
    losses = {}
    for f16 in (True, False):
        head = _head(E, 5376, 512, 5000, proto_f16=f16)
        w0 = head.w_proj.clone()
        loss = E.swav_train_step(gen, head, mean_latent, draws, cfg)
        losses[f16] = loss.item()
        assert math.isfinite(losses[f16]) and 0 < losses[f16] < 10 * math.log(5000)
        # rows of dS sum to zero -> so does the bias gradient (relative to its magnitude)
        gb = head.g_bias
        assert abs(gb.sum().item()) < 1e-3 * gb.abs().sum().item() + 1e-12
        for g in (head.g_proj, head.g_proto):
            assert torch.isfinite(g).all() and g.abs().max().item() > 0
        assert not torch.equal(head.w_proj, w0)           # the LARC/SGD step moved the weights
    # fp16x1 vs bf16x3 score GEMM at eps = 0.005 (scores x200 in the exponent): inside the stated 2e-3 loss tolerance
    assert abs(losses[True] - losses[False]) < 2e-3 * abs(losses[False]), losses


def test_generic_config_k8000_step(ffhq):
    """the generic hfc_with_swav_config.py (afhq-256 and every model without its own file): K = 8000 prototypes -
    the 512-thread Sinkhorn instantiation and the widest loss kernel, at the shipped eps / T"""
    E, gen, mean_latent = ffhq
    cfg = E.StepConfig(hlen=5376, patch_size=20000, num_patches=2, niters=10, eps=0.005, temperature=0.01,
                       truncation=0.7, perturb_std=[1.0] * 6)
    head = _head(E, 5376, 512, 8000)
    loss = E.swav_train_step(gen, head, mean_latent, _draws(E, 1, 512, 6, 65536, 2, 3), cfg)
    assert math.isfinite(loss.item()) and 0 < loss.item() < 10 * math.log(8000)
    gb = head.g_bias
    assert abs(gb.sum().item()) < 1e-3 * gb.abs().sum().item() + 1e-12
    assert torch.isfinite(head.g_proj).all() and torch.isfinite(head.g_proto).all()


@pytest.mark.parametrize("eps", [0.05, 0.005])
def test_ffhq256_sinkhorn_marginals_full_size(ffhq, eps):
    """Q = softmax_k(S/eps + log a) of a 40000 x 5000 score matrix: unit rows, uniform prototype marginals
    (eps = 0.005 is the shipped value: after 10 iterations the marginals are not converged there, as in the reference)"""
    E, gen, mean_latent = ffhq
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(0)
    n, k, c = 40000, 5000, 512
    z = torch.nn.functional.normalize(torch.randn(n, c, device="cuda"), dim=1)
    wk = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=1)
    zf, wf = L.round_f16(z), L.round_f16(wk)
    u0 = torch.zeros(k, device="cuda")
    s = L.gemm(zf, None, wf, None, n, k, c, 1, colexp=(u0, E.LOG2E / eps), pair=True)
    ws = L.SinkhornWorkspace(k, "cuda")
    la = E.sinkhorn_log_a(s, 10, eps, ws, n, u_first=u0)
    la_ref = E.sinkhorn_log_a(s, 10, eps, ws, n)                      # first pass by the streaming kernel
    torch.testing.assert_close(la, la_ref, rtol=0, atol=2e-4 if eps > 0.01 else 2e-3)
    # the training step's route: iterations 2.. through the 16-bit cache (gx_sinkhorn_pass_cached)
    la16 = E.sinkhorn_log_a(s, 10, eps, ws, n, u_first=u0, cache16=True)
    assert not torch.equal(la16, la)
    torch.testing.assert_close(la16, la, rtol=0, atol=1e-3)
    q16 = L.sinkhorn_q(s, 1.0 / eps, la16)
    q = L.sinkhorn_q(s, 1.0 / eps, la)
    big = q > 1e-9
    assert ((q16 - q).abs()[big] / q[big]).max().item() < 1e-3
    del q16, big
    torch.testing.assert_close(q.sum(1), torch.ones(n, device="cuda"), rtol=1e-4, atol=0)
    col = q.sum(0)
    # 10 iterations from random scores: marginals within a few percent of N/K, mean exactly N/K
    assert abs(col.mean().item() - n / k) < 1e-3 * n / k
    assert (col - n / k).abs().max().item() < (0.05 if eps > 0.01 else 0.5) * n / k
    assert (q >= 0).all()


def test_ffhq256_projection_routes_agree(ffhq):
    """per-resolution projection + upsample-sum == projection of gathered upsampled vectors (full D)"""
    E, gen, mean_latent = ffhq
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(5)
    lat = torch.randn(2, gen.n_latent, 512, device="cuda")
    _, feats = gen.synthesize(lat, None, need_image=False)
    wp = (torch.randn(512, 5376) / 5376 ** 0.5).cuda()
    wp_hi, wp_lo = L.split_planes(wp, want_lo=True)
    z_all, levels = E.project_all_pixels(wp_hi, wp_lo, feats, 2, 256, 256, 5376, 3)
    assert [lv["h"] for lv in levels] == [4, 8, 16, 32, 64, 128, 256]
    idx = torch.randint(0, 2 * 65536, (4096,), device="cuda", dtype=torch.int32)
    a_hi, a_lo, _ = L.gather_rows(feats, 256, 256, 5376, (idx // 65536).to(torch.int32), (idx % 65536).to(torch.int32),
                                  4096, want_lo=True)
    z_rows = L.gemm(a_hi, a_lo, wp_hi, wp_lo, 4096, 512, 5376, 3)
    ref = z_rows
    got = z_all[idx.long()]
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2e-4 * scale, ((got - ref).abs().max().item(), scale)


def test_car512_steps_and_label_map():
    """config 3: 512^2 generator, the 512^2-native maps fall outside hlen=5376, K=4000,
    eps=0.01, temperature=0.01 as shipped (hfc_with_swav_car_config.py:51-65; T/eps = 1 -> swav_loss_pow_kernel),
    6 perturbable layers; N=20000 as configured and the full-image Sinkhorn problem N=262144."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    gen = _generator(512)
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(512, 512, generator=torch.Generator().manual_seed(3)).cuda()).mean(0, keepdim=True)
    n_layers = 6
    for patch, npatch in ((20000, 5), (None, 1)):
        cfg = E.StepConfig(hlen=5376, patch_size=patch, num_patches=npatch, niters=10, eps=0.01, temperature=0.01,
                           truncation=0.7, perturb_std=[1.0] * n_layers)
        head = _head(E, 5376, 512, 4000)
        draws = _draws(E, 1, 512, n_layers, 512 * 512, npatch, 11)
        loss = E.swav_train_step(gen, head, mean_latent, draws, cfg)
        assert math.isfinite(loss.item()) and 0 < loss.item() < 10 * math.log(4000), loss.item()
        gb = head.g_bias
        assert abs(gb.sum().item()) < 1e-3 * gb.abs().sum().item() + 1e-12
        assert torch.isfinite(head.g_proj).all() and torch.isfinite(head.g_proto).all()
    # label map at 512^2
    w = torch.randn(1, 512, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        wl = gen.style(w.cuda())
    preds, labels = E.predict_codes(gen, head.w_proj, wl, mean_latent, 0.7, 5376)
    assert tuple(preds.shape) == (1, 512, 512, 512) and tuple(labels.shape) == (1, 512, 512)
    assert labels.dtype == torch.int64
    assert torch.equal(labels, preds.argmax(dim=1))


def test_pidray256_label_map_batch():
    """config 4: BagGAN channel map (sum C = 2528), projection 2528 -> 512, arg-max label maps, batch of 4"""
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.baggan import baggan_channels
    gen = _generator(256, seed=13, channels=baggan_channels())
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(512, 512, generator=torch.Generator().manual_seed(3)).cuda()).mean(0, keepdim=True)
        wl = gen.style(torch.randn(4, 512, generator=torch.Generator().manual_seed(9)).cuda())
    torch.manual_seed(2)
    wp = (torch.randn(512, 2528) / 2528 ** 0.5).cuda()
    preds, labels = E.predict_codes(gen, wp, wl, mean_latent, 0.9, 2528)
    assert tuple(labels.shape) == (4, 256, 256) and labels.dtype == torch.int64
    assert torch.equal(labels, preds.argmax(dim=1))
    # deterministic: the same call gives the same bits
    preds2, labels2 = E.predict_codes(gen, wp, wl, mean_latent, 0.9, 2528)
    assert torch.equal(labels, labels2) and torch.equal(preds, preds2)
    # images of the batch are independent: image 2 alone gives the same map
    _, l2 = E.predict_codes(gen, wp, wl[2:3], mean_latent, 0.9, 2528)
    assert torch.equal(l2[0], labels[2])


@pytest.mark.parametrize("tensor", [False, True, "gemm"])
@pytest.mark.parametrize("c,h,k", [(1024, 64, 32), (512, 128, 64), (1024, 8, 4)])
def test_kmeans_assign_full_shapes(c, h, k, tensor):
    """config 5: per-pixel nearest centre on [b, C, h, w] features, all routes: direct fp32 distances
    (SIMT), the fused tensor-core kernel (True: rows split in registers, mma.sync against centre fragments in shared
    memory, arg-min of ||c||^2 - 2 x.c in the epilogue) and the GEMM route ("gemm": 3-pass split-bf16 scores X C^T
    through gx_gemm, then gx_argmin_affine)"""
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(c + h)
    b = 2
    x = torch.randn(b * h * h, c, device="cuda") + 0.5          # non-centred features: x.c is not small against ||x||^2
    cen = torch.randn(k, c, device="cuda") + 0.5
    lab = L.kmeans_assign(x[:, :c // 2].contiguous(), cen, x[:, c // 2:].contiguous(), tensor=tensor)
    d = torch.cdist(x.double(), cen.double()) ** 2
    ref = d.argmin(1)
    mism = lab.long() != ref
    if mism.any():      # only where the two nearest centres are closer than the rounding of the route
        top2 = d.topk(2, dim=1, largest=False).values
        assert ((top2[:, 1] - top2[:, 0])[mism] < 1e-4 * top2[:, 0][mism]).all()
    assert mism.float().mean().item() < 1e-3
    if not tensor:      # the default route of a large labels-only call is the tensor-core one
        auto = L.kmeans_assign(x[:, :c // 2].contiguous(), cen, x[:, c // 2:].contiguous())
        want = L.kmeans_assign(x[:, :c // 2].contiguous(), cen, x[:, c // 2:].contiguous(), tensor=x.shape[0] >= 4096)
        assert torch.equal(auto, want)


def test_score_gemm_precision_modes_at_config_eps():
    """What the operand format of the pixel x prototype GEMM does to the codes at the shipped eps = 0.005
    (scores are multiplied by 200 inside the exponential): fp16 single pass (default) vs the 3-plane bf16
    split vs fp64, on the softmax_k(S/eps) of 4096 unit-norm rows x 5000 prototypes.  These are the stated
    tolerances of DESIGN.md §4.2."""
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(1)
    n, k, c, eps = 4096, 5000, 512, 0.005
    z = torch.nn.functional.normalize(torch.randn(n, c, device="cuda"), dim=1)
    wk = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=1)
    s64 = z.double() @ wk.double().t()
    s16 = L.gemm(L.round_f16(z), None, L.round_f16(wk), None, n, k, c, 1, pair=True)
    zh, zl = L.split_planes(z, want_lo=True)
    wh, wl = L.split_planes(wk, want_lo=True)
    s3 = L.gemm(zh, zl, wh, wl, n, k, c, 3, pair=True)
    e16, e3 = (s16.double() - s64).abs(), (s3.double() - s64).abs()
    print("|dS| fp16x1 max/rms", e16.max().item(), e16.pow(2).mean().sqrt().item(), " bf16x3", e3.max().item(),
          e3.pow(2).mean().sqrt().item())
    assert e16.max().item() < 1.5e-4 and e16.pow(2).mean().sqrt().item() < 2e-5
    assert e3.max().item() < 1e-5 and e3.pow(2).mean().sqrt().item() < 2e-6
    q64 = torch.softmax(s64 / eps, dim=1)
    big = q64 > 1e-4                                       # the entries that carry the assignment
    for s_, rms_tol, max_tol in ((s16, 6e-3, 4e-2), (s3, 6e-4, 4e-3)):
        q = torch.softmax(s_.double() / eps, dim=1)
        rel = ((q - q64).abs() / q64)[big]
        print("rel dQ rms/max", rel.pow(2).mean().sqrt().item(), rel.max().item())
        assert rel.pow(2).mean().sqrt().item() < rms_tol and rel.max().item() < max_tol, (rel.max().item(),)


def test_pretrain_and_evaluate_entry_points(tmp_path):
    """the reference's CLI surface (pretrain.py / evaluate.py arguments) on the ffhq-256 config: two
    optimiser steps, artefacts written like the reference's, then label maps from the saved head."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import evaluate
    import pretrain
    out = str(tmp_path / "run")
    swav = pretrain.main(["--model", "ffhq-256", "--out_dir", out, "--num_epochs", "2", "--num_test_samples", "1"])
    assert os.path.exists(os.path.join(out, "projection.pt")) and os.path.exists(os.path.join(out, "prototypes.pt"))
    proj = torch.load(os.path.join(out, "projection.pt"), weights_only=False)
    assert proj[0].weight.shape == (512, 5376) and torch.isfinite(proj[0].weight).all()
    res = evaluate.main(["--model", "ffhq-256", "--out_dir", out, "--num_test_samples", "2"])
    assert tuple(res["code_labels"].shape) == (2, 256, 256) and res["code_labels"].dtype == torch.int64
    assert os.path.exists(os.path.join(out, "tests", "labels.pt"))
    # same seed -> the same label maps (deterministic path)
    res2 = evaluate.main(["--model", "ffhq-256", "--out_dir", out, "--num_test_samples", "2"])
    assert torch.equal(res["code_labels"], res2["code_labels"])


def test_ffhq256_label_map_matches_reference_golden():
    """Full-size label map against the UNMODIFIED reference (tests/golden/labelmap_ffhq256.npz: predict_swav_codes of
    the reference on Generator(256, 512, 8) with seeded random-init weights, hlen 5376, 512 code channels):
    integer labels identical except pixels whose reference top-2 margin is inside the error band."""
    import os
    from oracle import ganecdotes_oracle as O
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "labelmap_ffhq256.npz"))
    g = {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}
    gen_seed, proj_seed, w_seed = [int(v) for v in g["seeds"]]
    sd = O.init_generator_state(256, 512, 8, gen_seed)
    gen = Generator(256, 512, 8)
    gen.load_state_dict(sd, strict=True)
    gen = gen.cuda()
    rg = torch.Generator().manual_seed(w_seed)
    torch.randn(256, 512, generator=rg)                 # the mean-latent draws come first
    z = torch.randn(2, 512, generator=rg)
    mean_latent = g["mean_latent"].cuda()
    w_proj = (torch.randn(512, 5376, generator=torch.Generator().manual_seed(proj_seed)) / 5376 ** 0.5).cuda()
    with torch.no_grad():
        w = gen.style(z.cuda())
    preds, labels = E.predict_codes(gen, w_proj, w, mean_latent, 0.7, 5376)
    assert tuple(preds.shape) == (2, 512, 256, 256) and labels.dtype == torch.int64
    for i in range(2):
        absmax = float(g[f"absmax{i}"])
        ref = g[f"preds{i}_sample"]
        assert (preds[i:i + 1, ::16, ::8, ::8].cpu() - ref).abs().max().item() < 5e-4 * absmax
        mism = labels[i].cpu() != g[f"labels{i}"][0].long()
        assert mism.float().mean().item() < 5e-3
        if mism.any():
            assert g[f"margin{i}"][0].float()[mism].max().item() < 1e-3 * absmax


def test_ffhq256_pretrain_step_matches_reference_golden():
    """BASELINE config 1 at full size against the UNMODIFIED reference (tests/golden/pretrain_ffhq256.npz: one
    CPU pretrain step of the reference, ffhq-256 geometry, recorded draws): loss within 2e-3, weight updates
    within 3 % (bf16 backward operands), like the small seeded run of test_pretrain_matches_reference_golden."""
    import os
    from oracle import ganecdotes_oracle as O
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "pretrain_ffhq256.npz"))
    g = {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}
    gen_seed, head_seed, _ = [int(v) for v in g["seeds"]]
    gen = Generator(256, 512, 8)
    gen.load_state_dict(O.init_generator_state(256, 512, 8, gen_seed), strict=True)
    gen = gen.cuda()
    hg = torch.Generator().manual_seed(head_seed)
    wp = torch.randn(512, 5376, generator=hg) / 5376 ** 0.5
    wk = torch.randn(5000, 512, generator=hg) / 512 ** 0.5
    bk = 0.01 * torch.randn(5000, generator=hg)
    head = E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01, 3, 1)

    def view(v):
        return E.ViewDraws(layer_no=[int(g[f"{v}_layer"])], pert_z=g[f"{v}_pert_z"].unsqueeze(0),
                           angle=[float(g[f"{v}_angle"])], flip=[bool(g[f"{v}_flip"])])
    draws = E.StepDraws(z=g["z"], view_s=view("s"), view_t=view("t"),
                        perms=[[g[f"perm{p}"].long()] for p in range(5)])
    cfg = E.StepConfig(hlen=5376, patch_size=20000, num_patches=5, niters=10, eps=0.005, temperature=0.01,
                       truncation=0.7, perturb_std=[1.0] * 6)
    loss = E.swav_train_step(gen, head, g["mean_latent"].cuda(), draws, cfg)
    ref_loss = float(g["losses"][0])
    assert abs(loss.item() - ref_loss) < 2e-3 * abs(ref_loss), (loss.item(), ref_loss)
    init = [wp, torch.nn.functional.normalize(wk, dim=1), bk]
    fin = [head.w_proj.cpu(), head.w_proto.cpu(), head.b_proto.cpu()]
    for f, i0, ref_norm in zip(fin, init, g["update_norms"].tolist()):
        assert abs((f - i0).norm().item() - ref_norm) < 3e-2 * ref_norm
    d_proj = (fin[0] - init[0])[::16, ::64]
    assert (d_proj - g["delta_w_proj_sample"]).norm().item() < 5e-2 * g["delta_w_proj_sample"].norm().item()


def test_car512_label_map_matches_reference_golden():
    """BASELINE config 3 geometry against the UNMODIFIED reference (tests/golden/labelmap_car512.npz): 512^2
    generator, 5504 feature channels sliced to hlen 5376 after upsampling."""
    import os
    from oracle import ganecdotes_oracle as O
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "labelmap_car512.npz"))
    g = {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}
    gen_seed, proj_seed, w_seed = [int(v) for v in g["seeds"]]
    gen = Generator(512, 512, 8)
    gen.load_state_dict(O.init_generator_state(512, 512, 8, gen_seed), strict=True)
    gen = gen.cuda()
    rg = torch.Generator().manual_seed(w_seed)
    torch.randn(256, 512, generator=rg)                 # the mean-latent draws come first
    z = torch.randn(2, 512, generator=rg)[:1]
    w_proj = (torch.randn(512, 5376, generator=torch.Generator().manual_seed(proj_seed)) / 5376 ** 0.5).cuda()
    with torch.no_grad():
        w = gen.style(z.cuda())
    preds, labels = E.predict_codes(gen, w_proj, w, g["mean_latent"].cuda(), 0.7, 5376)
    assert tuple(preds.shape) == (1, 512, 512, 512) and tuple(labels.shape) == (1, 512, 512)
    absmax = float(g["absmax0"])
    assert (preds[:, ::16, ::16, ::16].cpu() - g["preds0_sample"]).abs().max().item() < 5e-4 * absmax
    mism = (labels.cpu() != g["labels0"].long()).flatten()
    assert mism.float().mean().item() < 5e-3
    if mism.any():
        neartie = torch.from_numpy(np.unpackbits(g["neartie0"].numpy())[:mism.numel()].astype(bool))
        assert bool(neartie[mism].all())
