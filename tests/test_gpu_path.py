"""GPU parity tests of the assembled path: Generator drop-in, SwAV pretrain steps and
predict_swav_codes against the golden outputs of the unmodified reference, and the
batched (joint Sinkhorn) step against the CPU oracle."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ganecdotes_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


@pytest.fixture(scope="module")
def gen():
    from ganecdotes_b200.stylegan2.model import Generator
    sd = O.init_generator_state(16, 64, 2, 7)
    g = Generator(16, 64, 2)
    missing, unexpected = g.load_state_dict(sd, strict=True)
    return g.cuda()


def test_generator_state_dict_layout(gen):
    """same parameter / buffer names as the reference Generator (rosinality layout)"""
    sd = O.init_generator_state(16, 64, 2, 7)
    assert set(gen.state_dict().keys()) == set(sd.keys())


def test_generator_forward_matches_reference(gen):
    g = load("generator")
    w = gen.style(g["z"].cuda())
    torch.testing.assert_close(w.cpu(), g["w"], rtol=1e-4, atol=1e-4)
    img, feats = gen([g["z"].cuda()], truncation=0.7, truncation_latent=g["mean_latent"].cuda(),
                     input_is_latent=False, randomize_noise=False)
    assert [tuple(f.shape) for f in feats] == [(2, 512, 4, 4), (2, 512, 8, 8), (2, 512, 8, 8), (2, 512, 16, 16),
                                               (2, 512, 16, 16)]
    for i, f in enumerate(feats):
        ref = g[f"feat{i}"]
        got = f[:, ::16].cpu()
        scale = ref.abs().max().item()
        assert (got - ref).abs().max().item() < 5e-4 * scale, (i, (got - ref).abs().max().item(), scale)
        torch.testing.assert_close(f.double().sum(dim=(2, 3)).cpu(), g[f"feat{i}_sum"], rtol=1e-3, atol=5e-2)
    torch.testing.assert_close(img.cpu(), g["img"], rtol=1e-3, atol=2e-3 * g["img"].abs().max().item())
    img2, latent = gen([g["w"].cuda()], return_latents=True, truncation=0.7,
                       truncation_latent=g["mean_latent"].cuda(), input_is_latent=True, randomize_noise=False)
    torch.testing.assert_close(latent.cpu(), g["latent"], rtol=1e-5, atol=1e-5)
    # W+ input with explicit per-sample noise
    noises = [g[f"noise3_{i}"].cuda() for i in range(5)]
    img3, feats3 = gen([g["wplus"].cuda()], input_is_latent=True, noise=noises)
    for i, f in enumerate(feats3):
        ref = g[f"feat3_{i}"]
        assert (f[:, 5::32].cpu() - ref).abs().max().item() < 5e-4 * ref.abs().max().item()
    torch.testing.assert_close(img3.cpu(), g["img3"], rtol=1e-3, atol=2e-3 * g["img3"].abs().max().item())


def golden_cfg(g):
    hlen, nclasses, nproto, patch, npatch, nepochs, nl = [int(v) for v in g["cfg"]]
    cfg = dict(
        perturb_args=dict(truncation=0.7, n_layers=nl, n_samples=1, layer_no=None,
                          perturb_std=[float(v) for v in g["perturb_std"]]),
        swav_args=dict(num_epochs=nepochs, num_samples=1, num_patches=npatch, sampling_method='random',
                       patch_size=patch, hf_interp='nearest', warmup_epochs=nepochs, start_warmup=0.01,
                       use_scheduler=False, base_lr=0.01, final_lr=0.0001, trust_coeff=0.01,
                       freeze_prototype_niters=313, train_args=dict(lr=0.01, momentum=0.9),
                       projn_nw='linear', temperature=0.02, nprototypes=nproto, nclasses=nclasses,
                       hlen=hlen, add_local_loss=False, plot_test_images=False, epoch_print_freq=1,
                       max_masks=4),
        sinkhorn_args=dict(source_pdf='uniform', niters=10, eps=0.02),
        train=True, layer_hf_dim=[512, 1024, 1024])
    mc = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7, latent_dim=64, image_size=16)
    return cfg, mc


def test_pretrain_matches_reference_golden(gen, tmp_path):
    """Seeded end-to-end run of SwAVClustering.pretrain vs the seeded CPU run of the
    unmodified reference: same random stream, losses and weights within tolerance."""
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    g = load("swav")
    cfg, mc = golden_cfg(g)
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    torch.manual_seed(11)
    np.random.seed(11)
    obj = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=tb, **cfg)
    torch.testing.assert_close(obj.mean_latent.cpu(), g["mean_latent"], rtol=1e-4, atol=1e-5)
    recorded = []
    orig = obj.draw_step
    obj.draw_step = lambda b: recorded.append(orig(b)) or recorded[-1]
    init = {}
    from ganecdotes_b200.hfc_with_swav import engine as E
    orig_head = E.SwavHead

    def spy(*a, **k):
        init["p"] = [a[0].clone(), a[1].clone(), a[2].clone()]
        return orig_head(*a, **k)
    E.SwavHead = spy
    try:
        obj.pretrain(None, num_test_samples=0)
    finally:
        E.SwavHead = orig_head
    # identical random stream
    for e, d in enumerate(recorded):
        assert torch.equal(d.z, g[f"s{e}_z"])
        assert d.view_s.layer_no[0] == int(g[f"s{e}_s_layer"]) and d.view_t.layer_no[0] == int(g[f"s{e}_t_layer"])
        assert torch.equal(d.view_s.pert_z[0], g[f"s{e}_s_pert_z"])
        assert torch.equal(d.view_t.pert_z[0], g[f"s{e}_t_pert_z"])
        assert d.view_s.angle[0] == float(g[f"s{e}_s_angle"]) and d.view_t.flip[0] == bool(g[f"s{e}_t_flip"])
        assert torch.equal(d.perms[1][0], g[f"s{e}_perm1"])
    torch.testing.assert_close(init["p"][0].cpu(), g["init_w_proj"], rtol=0, atol=0)
    torch.testing.assert_close(init["p"][1].cpu(), g["init_w_proto"], rtol=0, atol=0)
    # losses: fp32 reference vs split-bf16 tensor-core path
    ref_losses = g["losses"].tolist()
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    # weights: compare the UPDATES (two LARC steps move each tensor by ~2e-4 relative)
    fin = [obj.projection[0].weight.data.cpu(), obj.prototype.weight.data.cpu(), obj.prototype.bias.data.cpu()]
    ref_fin = [g["final_w_proj"], g["final_w_proto"], g["final_b_proto"]]
    ref_init = [g["init_w_proj"], torch.nn.functional.normalize(g["init_w_proto"], dim=1), g["init_b_proto"]]
    for f, rf, ri in zip(fin, ref_fin, ref_init):
        upd = (rf - ri).norm().item()
        assert (f - rf).norm().item() < 3e-2 * upd, ((f - rf).norm().item(), upd)
    # artefacts: pickled modules like the reference
    assert os.path.exists(os.path.join(str(tmp_path), "projection.pt"))
    proj = torch.load(os.path.join(str(tmp_path), "projection.pt"), weights_only=False)
    assert isinstance(proj, torch.nn.Sequential) and proj[0].weight.shape == (64, 2560)

    # inference with the trained head: label map identical except near-ties
    with torch.no_grad():
        obj.projection[0].weight.data.copy_(g["final_w_proj"])
    preds, labels = obj.predict_swav_codes(g["pred_w"].cuda())
    assert tuple(preds.shape) == (1, 64, 16, 16) and labels.dtype == torch.int64 and tuple(labels.shape) == (1, 16, 16)
    ref_p = g["preds"]
    assert (preds[:, ::4].cpu() - ref_p).abs().max().item() < 5e-4 * ref_p.abs().max().item()
    mism = labels.cpu() != g["labels"]
    if mism.any():   # only allowed where the reference's own top-2 margin is inside our error band
        full_ref, _ = O.predict_codes(O.init_generator_state(16, 64, 2, 7), g["pred_w"], g["mean_latent"], 0.7,
                                      g["final_w_proj"], 2560)
        top2 = full_ref.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])[mism]
        assert margin.max().item() < 1e-3 * full_ref.abs().max().item()
    assert mism.float().mean().item() < 0.02


def test_pretrain_with_patch_sampling_matches_reference_golden(gen, tmp_path):
    """sampling_method='patch' (square crops at a random offset, ref swav_clustering.py:150-158, 383-385): seeded
    SwAVClustering.pretrain vs the seeded CPU run of the unmodified reference (tests/golden/swav_patch.npz)."""
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    g = load("swav_patch")
    base = load("swav")
    cfg, mc = golden_cfg(g)
    cfg["swav_args"]["sampling_method"] = 'patch'
    assert cfg["swav_args"]["patch_size"] == 10
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    torch.manual_seed(11)
    np.random.seed(11)
    obj = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=tb, **cfg)
    recorded = []
    orig = obj.draw_step
    obj.draw_step = lambda b: recorded.append(orig(b)) or recorded[-1]
    obj.pretrain(None, num_test_samples=0)
    from ganecdotes_b200.hfc_with_swav.engine import patch_pick_rows
    for e, d in enumerate(recorded):        # identical random stream, crops at the recorded offsets
        assert torch.equal(d.z, g[f"s{e}_z"])
        assert d.view_t.angle[0] == float(g[f"s{e}_t_angle"])
        for p in range(2):
            assert torch.equal(d.perms[p][0], patch_pick_rows(16, 16, int(g[f"s{e}_pick{p}"]), 10))
    ref_losses = g["losses"].tolist()
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    fin = [obj.projection[0].weight.data.cpu(), obj.prototype.weight.data.cpu(), obj.prototype.bias.data.cpu()]
    ref_fin = [g["final_w_proj"], g["final_w_proto"], g["final_b_proto"]]
    ref_init = [base["init_w_proj"], torch.nn.functional.normalize(g["init_w_proto"], dim=1), g["init_b_proto"]]
    for f, rf, ri in zip(fin, ref_fin, ref_init):
        upd = (rf - ri).norm().item()
        assert (f - rf).norm().item() < 3e-2 * upd, ((f - rf).norm().item(), upd)


def test_pretrain_with_one_layer_projection_matches_reference_golden(gen, tmp_path):
    """projn_nw='1-layer' (Linear without bias + LeakyReLU(0.01, inplace), ref swav_clustering.py:250-256): seeded
    SwAVClustering.pretrain and predict_swav_codes vs the seeded CPU run of the unmodified reference
    (tests/golden/swav_1layer.npz); the saved projection.pt is the reference's Sequential(Linear, LeakyReLU)."""
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    g = load("swav_1layer")
    base = load("swav")
    cfg, mc = golden_cfg(g)
    cfg["swav_args"]["projn_nw"] = '1-layer'
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    torch.manual_seed(11)
    np.random.seed(11)
    obj = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=tb, **cfg)
    recorded = []
    orig = obj.draw_step
    obj.draw_step = lambda b: recorded.append(orig(b)) or recorded[-1]
    obj.pretrain(None, num_test_samples=0)
    for e, d in enumerate(recorded):        # identical random stream
        assert torch.equal(d.z, g[f"s{e}_z"])
        assert torch.equal(d.perms[1][0], g[f"s{e}_perm1"])
    ref_losses = g["losses"].tolist()
    assert len(losses) == len(ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    fin = [obj.projection[0].weight.data.cpu(), obj.prototype.weight.data.cpu(), obj.prototype.bias.data.cpu()]
    ref_fin = [g["final_w_proj"], g["final_w_proto"], g["final_b_proto"]]
    ref_init = [base["init_w_proj"], torch.nn.functional.normalize(g["init_w_proto"], dim=1), g["init_b_proto"]]
    for f, rf, ri in zip(fin, ref_fin, ref_init):
        upd = (rf - ri).norm().item()
        assert (f - rf).norm().item() < 3e-2 * upd, ((f - rf).norm().item(), upd)
    proj = torch.load(os.path.join(str(tmp_path), "projection.pt"), weights_only=False)
    assert isinstance(proj, torch.nn.Sequential) and isinstance(proj[1], torch.nn.LeakyReLU)
    # inference through a re-loaded object (the slope comes from the pickled module)
    cfg["train"] = False
    obj2 = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=None, **cfg)
    obj2.mean_latent = obj.mean_latent          # the constructor drew a new one
    with torch.no_grad():
        obj2.projection[0].weight.data.copy_(g["final_w_proj"])
    preds, labels = obj2.predict_swav_codes(g["pred_w"].cuda())
    ref_p = g["preds"]
    assert (preds[:, ::8].cpu() - ref_p).abs().max().item() < 5e-4 * ref_p.abs().max().item()
    mism = labels.cpu() != g["labels"]
    if mism.any():   # only where the reference's own top-2 margin is inside the error band
        full_ref, _ = O.predict_codes(O.init_generator_state(16, 64, 2, 7), g["pred_w"], g["mean_latent"], 0.7,
                                      g["final_w_proj"], 2560, proj_slope=0.01)
        top2 = full_ref.topk(2, dim=1).values
        assert (top2[:, 0] - top2[:, 1])[mism].max().item() < 1e-3 * full_ref.abs().max().item()
    assert mism.float().mean().item() < 0.02


def make_draws(b, d, n_layers, hw, npatch, seed):
    from ganecdotes_b200.hfc_with_swav import engine as E
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)

    def view():
        return E.ViewDraws(layer_no=[int(rs.randint(n_layers)) for _ in range(b)],
                           pert_z=torch.randn(b, 2 * n_layers, d, generator=g),
                           angle=[float(rs.uniform(-10, 10)) for _ in range(b)],
                           flip=[bool(rs.rand() < 0.5) for _ in range(b)])
    return E.StepDraws(z=torch.randn(b, d, generator=g), view_s=view(), view_t=view(),
                       perms=[[torch.randperm(hw, generator=g) for _ in range(b)] for _ in range(npatch)])


def oracle_step(sd, mean_latent, draws, wp, wk, bk, hlen, patch, npatch, pstd, niters, eps, temp, bufs=None,
                mode="nearest", proj_slope=None):
    b = draws.z.shape[0]
    w = O.style_mlp(sd, draws.z)
    rows = {"s": [[] for _ in range(npatch)], "t": [[] for _ in range(npatch)]}
    for name, view in (("s", draws.view_s), ("t", draws.view_t)):
        for i in range(b):
            hf, _ = O.view_features(sd, w[i:i + 1], mean_latent, 0.7, view.layer_no[i], view.pert_z[i], 3, pstd, hlen,
                                    mode)
            hf = O.rotate_flip(hf, view.angle[i], view.flip[i])
            for p in range(npatch):
                rows[name][p].append(O.sample_rows(hf, draws.perms[p][i], patch))
    rs = [torch.cat(r) for r in rows["s"]]
    rt = [torch.cat(r) for r in rows["t"]]
    return O.swav_step(rs, rt, wp, wk, bk, niters, eps, temp, bufs, proj_slope=proj_slope)


@pytest.mark.parametrize("dedup,proto_f16,interp,slope", [(False, True, "nearest", None), (True, True, "nearest", None),
                                                            (True, False, "nearest", None),
                                                            (True, False, "bilinear", None),
                                                            (True, False, "nearest", 0.01),
                                                            (False, False, "nearest", 0.01),
                                                            (True, False, "bilinear", 0.01)])
def test_batched_joint_step_matches_oracle(gen, dedup, proto_f16, interp, slope):
    """B = 3 latents per step: joint-batch Sinkhorn over the row-concatenation.
    dedup=True: every pixel is projected once and the patches gather rows of Z (the path the
    full-size ffhq step takes, where 5 x 20000 samples > 65536 pixels).
    proto_f16: pixel x prototype scores from single fp16 planes of the unit-norm operands (default,
    |dS| ~ 1e-5) or from the 3-plane bf16 split; both meet the same tolerances.
    slope: projn_nw == '1-layer' (LeakyReLU after the projection, ref :250-256) on both projection routes."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200 import _lib as L
    sd = O.init_generator_state(16, 64, 2, 7)
    torch.manual_seed(0)
    hlen, c, k, patch, npatch = 2560, 64, 48, 120, 2
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    mean_latent = O.style_mlp(sd, torch.randn(64, 64)).mean(0, keepdim=True)
    pstd = [1.0, 0.5, 1.0]
    head = E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01, 3, 1,
                      proto_f16=proto_f16, proj_slope=slope)
    cfg = E.StepConfig(hlen=hlen, patch_size=patch, num_patches=npatch, niters=10, eps=0.02, temperature=0.02,
                       truncation=0.7, perturb_std=pstd, dedup=dedup, hf_interp=interp)
    bufs = None
    rwp, rwk, rbk = wp, wk, bk
    for step in range(2):
        draws = make_draws(3, 64, 3, 256, npatch, 100 + step)
        ref = oracle_step(sd, mean_latent, draws, rwp, rwk, rbk, hlen, patch, npatch, pstd, 10, 0.02, 0.02, bufs,
                          mode=interp, proj_slope=slope)
        loss = E.swav_train_step(gen, head, mean_latent.cuda(), draws, cfg)
        assert abs(loss.item() - ref["loss"].item()) < 2e-3 * abs(ref["loss"].item()), (loss.item(), ref["loss"])
        # gradients (bf16 backward GEMMs): relative Frobenius error
        for got, exp in zip((head.g_proj, head.g_proto, head.g_bias), ref["grads"]):
            rel = (got.cpu() - exp).norm().item() / exp.norm().item()
            assert rel < 2e-2, rel
        rwp, rwk, rbk = ref["params"]
        bufs = ref["bufs"]
        for got, exp in zip((head.w_proj, head.w_proto, head.b_proto), ref["params"]):
            torch.testing.assert_close(got.cpu(), exp, rtol=1e-4, atol=2e-6)


def test_full_precision_backward_option(gen):
    """passes_bwd = 3 tightens the gradients to the fp32 reference."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    sd = O.init_generator_state(16, 64, 2, 7)
    torch.manual_seed(1)
    hlen, c, k, patch, npatch = 2560, 64, 48, 200, 1
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    mean_latent = O.style_mlp(sd, torch.randn(64, 64)).mean(0, keepdim=True)
    pstd = [1.0, 1.0, 1.0]
    head = E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01, 3, 3)
    cfg = E.StepConfig(hlen=hlen, patch_size=patch, num_patches=npatch, niters=10, eps=0.02, temperature=0.02,
                       truncation=0.7, perturb_std=pstd)
    draws = make_draws(1, 64, 3, 256, npatch, 7)
    ref = oracle_step(sd, mean_latent, draws, wp, wk, bk, hlen, patch, npatch, pstd, 10, 0.02, 0.02)
    E.swav_train_step(gen, head, mean_latent.cuda(), draws, cfg)
    for got, exp in zip((head.g_proj, head.g_proto, head.g_bias), ref["grads"]):
        rel = (got.cpu() - exp).norm().item() / exp.norm().item()
        assert rel < 3e-3, rel


def test_api_parity_methods(gen, tmp_path):
    """sinkhorn_knopp / create_pixel_feature_vectors / get_swav_codes_from_hidden_features
    keep the reference's call signatures and results."""
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    g = load("swav")
    cfg, mc = golden_cfg(g)
    obj = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=None, **cfg)
    obj.eps = 0.005
    q = obj.sinkhorn_knopp(g["sk_scores_s"].cuda(), None)
    torch.testing.assert_close(q.cpu(), g["sk_q_s"], rtol=2e-3, atol=1e-9)
    sd = O.init_generator_state(16, 64, 2, 7)
    gg = load("generator")
    _, feats, _ = O.generator_forward(sd, gg["z"][:1], 0.7, gg["mean_latent"], False)
    hf_ref = O.pixel_feature_vectors(O.regroup_features(feats), 2560)
    hf = obj.create_pixel_feature_vectors([f.cuda() for f in O.regroup_features(feats)])
    assert torch.equal(hf.cpu(), hf_ref)
    obj.projection = torch.nn.Sequential(torch.nn.Linear(2560, 64, bias=False)).cuda()
    obj.prototype = torch.nn.Linear(64, 48).cuda()
    perm = torch.randperm(256)
    sc = obj.get_swav_codes_from_hidden_features(hf, picks=perm, train=True)
    ref = O.swav_scores(O.sample_rows(hf_ref, perm, 100), obj.projection[0].weight.data.cpu(),
                        obj.prototype.weight.data.cpu(), obj.prototype.bias.data.cpu())
    torch.testing.assert_close(sc.cpu(), ref, rtol=1e-3, atol=2e-4)
    codes = obj.get_swav_codes_from_hidden_features(hf, (1, 64, 16, 16), train=False)
    assert tuple(codes.shape) == (1, 64, 16, 16)


def test_image_source_pdf_step_matches_oracle(gen):
    """source_pdf == 'image' (cat config): Sinkhorn marginals from histograms of the per-pixel
    feature norm of the transformed tensor (ref :361-362, :523-532)."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    sd = O.init_generator_state(16, 64, 2, 7)
    torch.manual_seed(2)
    hlen, c, k, patch, npatch = 2560, 64, 48, 128, 1
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    mean_latent = O.style_mlp(sd, torch.randn(64, 64)).mean(0, keepdim=True)
    pstd = [1.0, 1.0, 1.0]
    draws = make_draws(1, 64, 3, 256, npatch, 21)
    w = O.style_mlp(sd, draws.z)
    rows, marg = {}, []
    for name, view in (("s", draws.view_s), ("t", draws.view_t)):
        hf, _ = O.view_features(sd, w, mean_latent, 0.7, view.layer_no[0], view.pert_z[0], 3, pstd, hlen)
        hf = O.rotate_flip(hf, view.angle[0], view.flip[0])
        rows[name] = [O.sample_rows(hf, draws.perms[0][0], patch)]
        marg.append(O.image_marginals(torch.norm(hf, p=2, dim=1), k, patch))
    ref = O.swav_step(rows["s"], rows["t"], wp, wk, bk, 10, 0.02, 0.02, marginals=(marg[0], marg[1]))
    head = E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01, 3, 3)
    cfg = E.StepConfig(hlen=hlen, patch_size=patch, num_patches=npatch, niters=10, eps=0.02, temperature=0.02,
                       truncation=0.7, perturb_std=pstd, source_pdf='image')
    loss = E.swav_train_step(gen, head, mean_latent.cuda(), draws, cfg)
    assert abs(loss.item() - ref["loss"].item()) < 3e-3 * abs(ref["loss"].item()), (loss.item(), ref["loss"])
    for got, exp in zip((head.g_proj, head.g_proto, head.g_bias), ref["grads"]):
        assert (got.cpu() - exp).norm().item() / exp.norm().item() < 2e-2


@pytest.mark.parametrize("size", [32])
def test_larger_generator_sliced_hlen(size):
    """car-512-like case at reduced size: more feature channels than hlen, the slice [:hlen]
    cuts through the concatenation; label map vs the oracle."""
    from ganecdotes_b200.stylegan2.model import Generator
    from ganecdotes_b200.hfc_with_swav import engine as E
    sd = O.init_generator_state(size, 64, 2, 9)
    g = Generator(size, 64, 2)
    g.load_state_dict(sd, strict=True)
    g = g.cuda()
    torch.manual_seed(0)
    total = 512 * (1 + 2 * 3)
    hlen = total - 304                      # not on a map boundary (TMA needs hlen % 8 == 0)
    wp = torch.randn(48, hlen) / hlen ** 0.5
    mean_latent = O.style_mlp(sd, torch.randn(32, 64)).mean(0, keepdim=True)
    w = O.style_mlp(sd, torch.randn(2, 64))
    ref_p, ref_l = O.predict_codes(sd, w, mean_latent, 0.7, wp, hlen)
    preds, labels = E.predict_codes(g, wp.cuda(), w.cuda(), mean_latent.cuda(), 0.7, hlen)
    assert (preds.cpu() - ref_p).abs().max().item() < 5e-4 * ref_p.abs().max().item()
    mism = labels.cpu() != ref_l
    if mism.any():
        top2 = ref_p.topk(2, dim=1).values
        assert (top2[:, 0] - top2[:, 1])[mism].max().item() < 1e-3 * ref_p.abs().max().item()
    assert mism.float().mean().item() < 0.02


def test_baggan_generator_and_label_map():
    """pidray-256 config: BagGAN StyleGANGenerator state dict -> drop-in Generator
    (key mapping + narrow channel map), features vs the reference golden, label map vs the oracle."""
    from ganecdotes_b200.baggan import generator_from_baggan, baggan_channels
    from ganecdotes_b200.hfc_with_swav import engine as E
    g = load("baggan")
    size = int(g["size"])
    sd = O.init_generator_state(size, 512, 8, 13, channels=O.baggan_channels())
    baggan_sd = {O.to_baggan_key(k): v for k, v in sd.items()}
    baggan_sd["head_m.0.weight"] = torch.zeros(1, 1, 3, 3)          # present in real checkpoints, unused
    gen = generator_from_baggan(baggan_sd, img_resolution=size)
    assert {r: c for r, c in baggan_channels().items() if r <= 256} == {r: c for r, c in O.baggan_channels().items()
                                                                        if r <= 256}
    img, feats = gen([g["w"].cuda()], truncation=0.9, truncation_latent=g["mean_latent"].cuda(),
                     input_is_latent=True, randomize_noise=False)
    assert [f.shape[1] for f in feats] == g["chans"].tolist()
    for i, f in enumerate(feats):
        cs = 16 if f.shape[1] >= 128 else 4
        sp = 1 if f.shape[-1] <= 16 else (2 if f.shape[-1] <= 32 else f.shape[-1] // 16)
        ref = g[f"feat{i}"]
        err = (f[:, 1::cs, ::sp, ::sp].cpu() - ref).abs().max().item()
        assert err < 1e-3 * ref.abs().max().item(), (i, err, ref.abs().max().item())
    ref_img = g["img"]
    assert (img[:, :, ::4, ::4].cpu() - ref_img).abs().max().item() < 2e-3 * ref_img.abs().max().item()
    # one-shot inference head: projection 2528 -> 512, arg-max label map (evaluate.py path)
    torch.manual_seed(0)
    hlen = 2528
    wp = torch.randn(512, hlen) / hlen ** 0.5
    w1 = g["w"][:1]
    ref_p, ref_l = O.predict_codes(sd, w1, g["mean_latent"], 0.9, wp, hlen)
    preds, labels = E.predict_codes(gen, wp.cuda(), w1.cuda(), g["mean_latent"].cuda(), 0.9, hlen)
    assert labels.shape == (1, size, size) and labels.dtype == torch.int64
    mism = labels.cpu() != ref_l
    if mism.any():
        top2 = ref_p.topk(2, dim=1).values
        assert (top2[:, 0] - top2[:, 1])[mism].max().item() < 2e-3 * ref_p.abs().max().item()
    assert mism.float().mean().item() < 0.02, mism.float().mean().item()


@pytest.mark.parametrize("size", ["XXS", "XS", "S"])
def test_one_shot_segmentor_head_matches_reference(size):
    """SURVEY §8(f) rank 1 (inference half): drop-in OneShotSegmentor on the tcgen05 conv kernel against the
    outputs of the unmodified reference module (golden) - class scores and the arg-max label map."""
    from ganecdotes_b200.hfc_with_swav import OneShotSegmentor
    g = np.load(os.path.join(GOLD, "segmentor.npz"))
    n_class = int(g[f"{size}_nclass"])
    torch.manual_seed(100 + n_class)
    net = OneShotSegmentor(512, n_class, size=size)                         # same default init, same RNG order
    assert list(net.state_dict().keys()) == [str(k) for k in g[f"{size}_keys"]]
    for v, ref_sum in zip(net.state_dict().values(), g[f"{size}_param_sums"]):
        assert abs(float(v.double().sum()) - float(ref_sum)) < 1e-9
    net = net.cuda().eval()
    x = torch.randn(2, 512, 24, 24, generator=torch.Generator().manual_seed(21)).cuda()
    ref = torch.from_numpy(g[f"{size}_y"])
    with torch.no_grad():
        y = net(x)
        labels = net.predict_labels(x.contiguous(memory_format=torch.channels_last))
        from ganecdotes_b200 import _lib as L
        planes = L.split_planes(x.permute(0, 2, 3, 1).reshape(-1, 512).contiguous(), want_lo=True)
        assert torch.equal(labels, net.predict_labels(x, planes))          # operand planes handed over by the producer
    assert y.shape == ref.shape
    scale = ref.abs().max().item()
    assert (y.cpu() - ref).abs().max().item() < 2e-4 * scale
    ref_l = torch.from_numpy(g[f"{size}_labels"])
    assert labels.dtype == torch.int64 and labels.shape == ref_l.shape
    mism = labels.cpu() != ref_l
    if mism.any():      # only where the reference's own top-2 margin is inside the error band
        top2 = ref.topk(2, dim=1).values
        assert (top2[:, 0] - top2[:, 1])[mism].max().item() < 4e-4 * scale
    assert mism.float().mean().item() < 0.01


@pytest.mark.parametrize("size", ["XXS", "XS", "S"])
def test_one_shot_segmentor_finetune_gradients(size):
    """SURVEY §8(f) rank 1, second half: one fine-tune step's gradients (CE loss through the segmentor,
    src/one_shot_pipeline.py:559-570) against torch autograd of the same conv stack in fp64 (F.conv2d)."""
    from ganecdotes_b200.hfc_with_swav import OneShotSegmentor
    n_class = 6
    torch.manual_seed(5)
    net = OneShotSegmentor(512, n_class, size=size).cuda().train()
    x = torch.randn(2, 512, 20, 28, generator=torch.Generator().manual_seed(8)).cuda()
    ref_state = {k: v.detach().double().cpu().requires_grad_(True) for k, v in net.state_dict().items()}
    y_ref = O.one_shot_segmentor(ref_state, x.double().cpu(), n_class, size)
    labels = torch.randint(0, y_ref.shape[1], (2, 20, 28), generator=torch.Generator().manual_seed(9))
    loss_ref = torch.nn.functional.cross_entropy(y_ref, labels)
    loss_ref.backward()
    y = net(x)
    assert y.requires_grad and y.shape == y_ref.shape
    loss = torch.nn.functional.cross_entropy(y, labels.cuda())
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    loss.backward()
    for name, p in net.named_parameters():
        gref = ref_state[name].grad.float()
        rel = (p.grad.cpu() - gref).norm().item() / gref.norm().item()
        assert rel < 2e-3, (name, rel)
    # an Adam step through the public module works end to end
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt.step()
    with torch.no_grad():
        loss2 = torch.nn.functional.cross_entropy(net.eval()(x), labels.cuda())
    assert loss2.item() < loss.item()


def test_create_hidden_features_from_perturbed_vectors_matches_oracle(gen, tmp_path):
    """SwAVClustering.create_hidden_features_from_perturbed_vectors (ref swav_clustering.py:574-656) and the
    lib/oneshot drop-ins (ref image_augmentor.py:8-104) against the oracle's view_features on the same draws."""
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    from ganecdotes_b200 import oneshot
    g = load("swav")
    cfg, mc = golden_cfg(g)
    obj = SwAVClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=None, **cfg)
    obj.match_reference_rng = False
    sd = O.init_generator_state(16, 64, 2, 7)
    mean_latent = obj.mean_latent.cpu()
    nl, pstd = cfg["perturb_args"]["n_layers"], cfg["perturb_args"]["perturb_std"]
    torch.manual_seed(11)
    z = torch.randn(1, 64)
    w = O.style_mlp(sd, z)
    for layer_no in (0, nl - 1):
        torch.manual_seed(100 + layer_no)
        pert_z = torch.cat([torch.randn(1, 64) for _ in range(2 * nl)], 0)      # the draws _draw_view makes
        hf_ref, img_ref = O.view_features(sd, w, mean_latent, 0.7, layer_no, pert_z, nl, pstd, cfg["swav_args"]["hlen"])
        torch.manual_seed(100 + layer_no)
        hf, img, l = obj.create_hidden_features_from_perturbed_vectors(layer_no=layer_no, input_latent=w.cuda())
        assert l == layer_no and tuple(hf.shape) == tuple(hf_ref.shape)
        scale = hf_ref.abs().max().item()
        assert (hf.cpu() - hf_ref).abs().max().item() < 5e-4 * scale
        assert (img.cpu() - img_ref).abs().max().item() < 5e-4 * max(1.0, img_ref.abs().max().item())
        # the same view through the lib/oneshot functions, called the way the reference calls them (:617-650)
        wt = mean_latent + 0.7 * (w - mean_latent)
        wplus = wt.unsqueeze(1).repeat(1, gen.n_latent, 1)
        stds = [0] * (2 * nl)
        stds[2 * layer_no] = stds[2 * layer_no + 1] = pstd[layer_no]
        torch.manual_seed(100 + layer_no)
        pl = oneshot.create_perturbed_vectors_from_latents(wplus, gen, n_samples=1, n_layers=nl, perturb_std=stds)
        assert len(pl) == 2 * nl and all(tuple(p.shape) == (1, 64) for p in pl)
        new = wplus.clone()
        new[:, 2 * layer_no, :], new[:, 2 * layer_no + 1, :] = pl[2 * layer_no], pl[2 * layer_no + 1]
        ref_new = O.perturbed_wplus(sd, w, mean_latent, 0.7, layer_no, pert_z, nl, pstd)
        torch.testing.assert_close(new, ref_new, rtol=1e-4, atol=1e-5)
        imgs, feats = oneshot.create_images_and_features_from_perturbed_latents(
            new.cuda(), gen, {'truncation': 0.7, 'mean_latent': obj.mean_latent})
        assert len(feats) == 1 + gen.num_layers // 2        # 13 -> 7 regrouping (5 -> 3 for the 16^2 generator)
        hf2 = obj.create_pixel_feature_vectors(feats)
        assert (hf2.cpu() - hf_ref).abs().max().item() < 5e-4 * scale
        only = oneshot.create_images_and_features_from_perturbed_latents(
            new.cuda(), gen, {'truncation': 0.7, 'mean_latent': obj.mean_latent}, layer_no=1, return_image=False)
        assert torch.equal(only, feats[1])
        skip = oneshot.create_images_and_features_from_perturbed_latents(
            new.cuda(), gen, {'truncation': 0.7, 'mean_latent': obj.mean_latent}, return_image=False, skip_const=True)
        assert len(skip) == len(feats) - 1 and torch.equal(skip[0], feats[1])


# ----------------------------------------------------------------------------------------
# SimCLR baseline head
# ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("n2,c", [(12, 32), (40, 512)])
def test_simclr_head_kernels_vs_oracle_autograd(n2, c):
    """BatchNorm1d + LeakyReLU forward / backward with the folded row scale, and the contrastive loss + gradient
    (gx_bn_stats / gx_bn_act_apply / gx_bn_act_bwd / gx_simclr_loss) against autograd of the oracle's formulas"""
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(n2 + c)
    hraw = torch.randn(n2, c) * 3
    rs = torch.rand(n2) + 0.5
    gamma, beta = torch.rand(c) + 0.5, 0.1 * torch.randn(c)
    da = torch.randn(n2, c)
    hr = hraw.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    h = hr * rs[:, None]
    mean, var = h.mean(0), h.var(0, unbiased=False)
    a_ref = torch.nn.functional.leaky_relu((h - mean) / torch.sqrt(var + 1e-5) * gr + br, 0.01)
    a_ref.backward(da)
    run_m, run_v = torch.zeros(c).cuda(), torch.ones(c).cuda()
    m, isd = L.bn_stats(hraw.cuda(), rs.cuda(), 1e-5, run_m, run_v, 0.1)
    torch.testing.assert_close(m.cpu(), mean.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(run_v.cpu(), 0.9 + 0.1 * h.detach().var(0, unbiased=True), rtol=1e-5, atol=1e-6)
    a, a_hi, a_lo = L.bn_act_apply(hraw.cuda(), rs.cuda(), m, isd, gamma.cuda(), beta.cuda(), 0.01)
    torch.testing.assert_close(a.cpu(), a_ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close((a_hi.float() + a_lo.float()).cpu(), a_ref.detach(), rtol=1e-4, atol=1e-4)
    dhs, dg, db = L.bn_act_bwd(da.cuda(), hraw.cuda(), rs.cuda(), m, isd, gamma.cuda(), beta.cuda(), 0.01)
    torch.testing.assert_close(dhs.cpu(), hr.grad, rtol=2e-3, atol=2e-5)
    torch.testing.assert_close(dg.cpu(), gr.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(db.cpu(), br.grad, rtol=1e-3, atol=1e-4)
    for temp in (1.0, 0.5):
        z = torch.randn(n2, c, requires_grad=True)
        ref = O.simclr_loss(z, temp)
        ref.backward()
        loss, dz = L.simclr_loss(z.detach().cuda(), temp)
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
        torch.testing.assert_close(dz.cpu(), z.grad, rtol=2e-3, atol=1e-6)


def test_simclr_pretrain_and_codes_match_reference_golden(gen, tmp_path):
    """SimCLRClustering on the recorded draws of the seeded run of the UNMODIFIED reference (tests/golden/simclr.npz):
    per-iteration losses, final weights + BatchNorm running statistics, eval-mode codes and label map."""
    from ganecdotes_b200.hfc_with_simclr import SimCLRClustering, SimCLRHead, simclr_train_step
    from ganecdotes_b200.hfc_with_simclr.simclr_clustering import SimCLRDraws
    g = load("simclr")
    hlen, nclasses, batch, niters, nl = [int(v) for v in g["cfg"]]
    pstd = [float(v) for v in g["perturb_std"]]
    mc = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7, latent_dim=64, image_size=16)
    cfg = dict(perturb_args=dict(truncation=0.7, n_layers=nl, n_samples=1, layer_no=None, perturb_std=pstd),
               simclr_args=dict(num_iters=niters, batch_size=batch, patch_size=100, hf_interp='nearest', trust_coeff=0.01,
                                train_args=dict(lr=0.01, momentum=0.9), temperature=1.0, nclasses=nclasses, hlen=hlen,
                                epoch_print_freq=1, max_masks=4), train=True, layer_hf_dim=[512, 1024, 1024])
    obj = SimCLRClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=None, **cfg)
    obj.mean_latent = g["mean_latent"].cuda()
    proj = torch.nn.Sequential(torch.nn.Linear(hlen, nclasses, bias=False), torch.nn.BatchNorm1d(nclasses),
                               torch.nn.LeakyReLU(inplace=True), torch.nn.Linear(nclasses, nclasses, bias=False)).cuda()
    with torch.no_grad():
        proj[0].weight.copy_(g["init_w1"]); proj[1].weight.copy_(g["init_bn_w"])
        proj[1].bias.copy_(g["init_bn_b"]); proj[3].weight.copy_(g["init_w2"])
    head = SimCLRHead(proj, 0.01, 0.9, 0.01)
    for e in range(niters):
        draws = SimCLRDraws(z=g[f"s{e}_z"], layer_no=[int(g[f"s{e}_s_layer"]), int(g[f"s{e}_t_layer"])],
                            pert_z=torch.stack([g[f"s{e}_s_pert_z"], g[f"s{e}_t_pert_z"]]),
                            angle=[float(g[f"s{e}_s_angle"]), float(g[f"s{e}_t_angle"])],
                            flip=[bool(g[f"s{e}_s_flip"]), bool(g[f"s{e}_t_flip"])], perm=g[f"s{e}_perm"].long())
        loss, _ = simclr_train_step(gen, head, obj.mean_latent, draws, hlen, batch, 1.0, 0.7, pstd)
        assert abs(loss.item() - float(g["losses"][e])) < 2e-3 * abs(float(g["losses"][e])), (e, loss.item())
    for got, key in zip(head.params, ("final_w1", "final_bn_w", "final_bn_b", "final_w2")):
        ref = g[key]
        upd = (ref - g[key.replace("final", "init")]).norm().item()
        assert (got.cpu() - ref).norm().item() < 5e-2 * upd + 1e-7, key          # the UPDATE agrees within 5 %
    torch.testing.assert_close(proj[1].running_mean.cpu(), g["final_bn_mean"], rtol=2e-3, atol=1e-5)
    torch.testing.assert_close(proj[1].running_var.cpu(), g["final_bn_var"], rtol=2e-3, atol=1e-5)
    obj.projection = proj.eval()
    preds, labels = obj.predict_simclr_codes(g["pred_w"].cuda())
    assert tuple(preds.shape) == tuple(g["preds"].shape) and labels.dtype == torch.int64
    scale = g["preds"].abs().max().item()
    assert (preds.cpu() - g["preds"]).abs().max().item() < 5e-3 * scale
    assert (labels.cpu() != g["labels"]).float().mean().item() < 0.02
    # API: pretrain() from the CPU random stream writes the reference's artefact
    obj2 = SimCLRClustering(gen, mc, out_dir=str(tmp_path), device='cuda', tb=None, **cfg)
    obj2.pretrain(None)
    assert os.path.exists(os.path.join(str(tmp_path), "projection.pt"))
    p2, l2 = obj2.predict_simclr_codes(g["pred_w"].cuda())
    assert torch.isfinite(p2).all() and tuple(l2.shape) == (1, 16, 16)


def test_kmeans_preprocessor_train_and_predict(gen, tmp_path):
    """HFCPreprocessor (ref baseline/hfc_kmeans/segmentor.py:11-226): fit on the latent-perturbed samples of one
    latent, one-hot cluster maps in {-1, +1}; the labels are the nearest fitted centre of the oracle's assignment on
    the same hidden features; artefact round trip."""
    from ganecdotes_b200.hfc_kmeans import HFCPreprocessor, preprocessor
    from ganecdotes_b200 import oneshot
    assert preprocessor is HFCPreprocessor
    mc = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7, latent_dim=64, image_size=16)
    cfg = dict(perturb_args=dict(truncation=0.7, n_layers=2, n_samples=4, perturb_std=[1.0, 1.0]), hfc_algo='hfc_kmeans',
               hfc_args=dict(kmeans_args=dict(verbose=0),
                             base_args=dict(out_dir=None, n_layers=2, clusters_per_layer=[4, 8], out_size=16,
                                            presaved=False)),
               hier_encode=False, hle_samples=10)
    torch.manual_seed(21)
    pre = HFCPreprocessor(model=gen, model_config=mc, out_dir=str(tmp_path), logger=None, train=True, **cfg)
    w = gen.style(torch.randn(1, 64).cuda())
    hidden, new_latents = pre.train_hfc_model(w, return_aug=True)
    assert [tuple(h.shape) for h in hidden] == [(4, 1024, 8, 8), (4, 1024, 16, 16)] and len(new_latents) == 4
    assert [tuple(c.shape) for c in pre.hfc_model.centers] == [(4, 1024), (8, 1024)]
    torch.manual_seed(22)
    preds, labels = pre.predict_hfc_vectors(w)
    assert tuple(preds.shape) == (1, 12, 16, 16) and set(preds.unique().tolist()) <= {-1.0, 1.0}
    assert (preds > 0).sum(1).eq(2).all()                         # one active cluster per layer and pixel
    # the same hidden features through the oracle's assignment with the fitted centres
    torch.manual_seed(22)
    mean_latent = gen.mean_latent(64)
    _, wl = gen([w], return_latents=True, truncation_latent=mean_latent, truncation=0.7, input_is_latent=True)
    _, hf = oneshot.create_images_and_features_from_perturbed_latents(wl, gen, {'truncation': 0.7, 'mean_latent': mean_latent},
                                                                      skip_const=True)
    for n in range(2):
        ref = O.kmeans_assign(hf[n].permute(0, 2, 3, 1).reshape(-1, 1024).cpu(), pre.hfc_model.centers[n].cpu())
        assert torch.equal(labels[n].flatten().cpu(), ref)
    # artefact: a second preprocessor in eval mode loads the centres and gives the same maps
    pre2 = HFCPreprocessor(model=gen, model_config=mc, out_dir=str(tmp_path), logger=None, train=False, **cfg)
    torch.manual_seed(22)
    preds2, _ = pre2.predict_hfc_vectors(w)
    assert torch.equal(preds, preds2)
    with pytest.raises(NotImplementedError):
        HFCPreprocessor(model=gen, model_config=mc, out_dir=str(tmp_path), **dict(cfg, hier_encode=True))


def test_cli_entry_points_for_the_baselines(tmp_path):
    """pretrain.py / evaluate.py with --method hfc_with_simclr and hfc_kmeans (the reference's CLI accepts the three
    methods, pretrain.py:45-57): artefacts written, label maps produced"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import evaluate
    import pretrain
    out = str(tmp_path / "simclr")
    pretrain.main(["--model", "ffhq-256", "--method", "hfc_with_simclr", "--out_dir", out, "--num_epochs", "2"])
    assert os.path.exists(os.path.join(out, "projection.pt"))
    res = evaluate.main(["--model", "ffhq-256", "--method", "hfc_with_simclr", "--out_dir", out, "--num_test_samples", "1"])
    assert tuple(res["code_labels"].shape) == (1, 256, 256)
    out = str(tmp_path / "kmeans")
    pretrain.main(["--model", "ffhq-256", "--method", "hfc_kmeans", "--out_dir", out])
    assert os.path.exists(os.path.join(out, "kmeans_centers.pt"))
    res = evaluate.main(["--model", "ffhq-256", "--method", "hfc_kmeans", "--out_dir", out, "--num_test_samples", "2"])
    assert [tuple(l.shape) for l in res["layer_labels"]] == [(2, 1, 8 * 2 ** n, 8 * 2 ** n) for n in range(5)]


def test_predict_graph_replay_is_bit_identical_to_predict_codes(gen):
    """engine.PredictGraph: the label-map path of a fixed batch captured as a CUDA graph; replays on new latents give
    the bits of the stream-launched predict_codes"""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(9)
    hlen = 2560
    wp = (torch.randn(64, hlen) / hlen ** 0.5).cuda()
    mean_latent = gen.style(torch.randn(64, 64).cuda()).mean(0, keepdim=True)
    g = E.PredictGraph(gen, wp, mean_latent, 0.7, hlen, batch=3)
    for seed in (1, 2, 3):
        w = gen.style(torch.randn(3, 64, generator=torch.Generator().manual_seed(seed)).cuda())
        ref_p, ref_l = E.predict_codes(gen, wp, w, mean_latent, 0.7, hlen)
        got_p, got_l = g(w)
        assert torch.equal(got_l, ref_l) and torch.equal(got_p, ref_p)
    with pytest.raises(ValueError):
        g(torch.zeros(2, 64, device="cuda"))


def test_train_graph_replay_matches_eager_steps(gen):
    """engine.TrainGraph: the optimiser step captured as a CUDA graph - two replays on new inputs give the losses and
    weights of two stream-launched steps from the same state (split-K accumulation order is the only difference)"""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(2)
    hlen, c, k = 2560, 64, 48
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    mean_latent = gen.style(torch.randn(64, 64).cuda()).mean(0, keepdim=True)
    cfg = E.StepConfig(hlen=hlen, patch_size=96, num_patches=2, niters=10, eps=0.02, temperature=0.02, truncation=0.7,
                       perturb_std=[1.0] * 3)
    draws = [make_draws(2, 64, 3, 256, 2, 30 + i) for i in range(3)]
    heads = [E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01, 3, 3) for _ in range(2)]
    inps = [[E.prepare_step_inputs(gen, d, cfg, "cuda") for d in draws] for _ in range(2)]
    ref = [E.swav_train_step_device(gen, heads[0], mean_latent, inps[0][i], cfg).item() for i in range(3)]
    first = E.swav_train_step_device(gen, heads[1], mean_latent, inps[1][0], cfg).item()       # eager warm-up step
    g = E.TrainGraph(gen, heads[1], mean_latent, inps[1][1], cfg)
    got = [first] + [g(inps[1][i]).item() for i in (1, 2)]
    for a, b_ in zip(got, ref):
        assert abs(a - b_) < 1e-5 * abs(b_), (got, ref)
    for a, b_ in zip((heads[1].w_proj, heads[1].w_proto, heads[1].b_proto), (heads[0].w_proj, heads[0].w_proto, heads[0].b_proto)):
        torch.testing.assert_close(a, b_, rtol=1e-5, atol=1e-7)
    assert heads[1].steps == 3
    with pytest.raises(RuntimeError):
        E.TrainGraph(gen, E.SwavHead(wp.clone().cuda(), wk.clone().cuda(), bk.clone().cuda(), 0.01, 0.9, 0.01), mean_latent,
                     inps[1][0], cfg)
