"""CPU tests of the host-side logic around the kernels (no device): resolution grouping of the feature
pyramid, split-K choice, the all-pixel (dedup) decision, segmentor layer construction."""
import types

import pytest
import torch

from ganecdotes_b200.hfc_with_swav import engine as E


def _feats(size, chans):
    """NHWC feature list of a StyleGAN2 pyramid: one map at 4x4, two per further resolution"""
    out, r = [], 4
    out.append(torch.zeros(1, r, r, chans[0]))
    for c in chans[1:]:
        r *= 2
        out += [torch.zeros(1, r, r, c), torch.zeros(1, r, r, c)]
    assert r == size
    return out


def test_resolution_groups_ffhq_and_sliced():
    feats = _feats(256, [512, 512, 512, 512, 512, 256, 128])
    g = E.resolution_groups(feats, 5376)
    assert [(x["h"], x["keep"], x["off"]) for x in g] == [(4, 512, 0), (8, 1024, 512), (16, 1024, 1536),
                                                          (32, 1024, 2560), (64, 1024, 3584), (128, 512, 4608),
                                                          (256, 256, 5120)]
    assert sum(x["keep"] for x in g) == 5376 and all(len(x["maps"]) in (1, 2) for x in g)
    # car-512: the 512^2 maps fall outside hlen; a cut inside a resolution keeps a partial map
    feats = _feats(512, [512, 512, 512, 512, 512, 256, 128, 64])
    g = E.resolution_groups(feats, 5376)
    assert g[-1]["h"] == 256 and sum(x["keep"] for x in g) == 5376
    g = E.resolution_groups(feats, 5376 - 200)
    assert g[-1]["h"] == 256 and g[-1]["keep"] == 56 and len(g[-1]["maps"]) == 1
    g = E.resolution_groups(feats, 5376 + 72)
    assert g[-1]["h"] == 512 and g[-1]["keep"] == 72 and len(g[-1]["maps"]) == 2


def test_pick_split_k_fills_the_grid():
    for tiles, kit in [(40, 2500), (20, 313), (2, 79), (300, 100)]:
        s = E.pick_split_k(tiles, kit, 148)
        assert 1 <= s <= max(1, kit // 8)
        work = tiles * s
        eff = work / (148 * -(-work // 148))
        assert eff >= 0.5 or s == max(1, min(64, kit // 8))       # at least half of the last wave busy


def test_all_pixel_projection_decision():
    cfg = types.SimpleNamespace(dedup=None, patch_size=20000, num_patches=5)
    assert E.use_dedup(cfg, 256, 256)                      # ffhq: 5 x 20000 samples over 65536 pixels
    cfg = types.SimpleNamespace(dedup=None, patch_size=1000, num_patches=2)
    assert not E.use_dedup(cfg, 256, 256)                  # a small fraction of the image: gather rows instead
    cfg = types.SimpleNamespace(dedup=None, patch_size=None, num_patches=1)
    assert E.use_dedup(cfg, 512, 512)                      # full-image Sinkhorn problem
    cfg = types.SimpleNamespace(dedup=False, patch_size=None, num_patches=1)
    assert not E.use_dedup(cfg, 512, 512)                  # explicit override


@pytest.mark.parametrize("size,n_class,expect", [
    ("XXS", 7, [(512, 12, 1)]),                                           # quirk 11: n_class is ignored
    ("XS", 7, [(512, 16, 1), (16, 8, 2), (8, 7, 1)]),
    ("S", 5, [(512, 128, 1), (128, 64, 2), (64, 64, 1), (64, 32, 2), (32, 5, 1)]),
])
def test_segmentor_layer_stack_matches_reference_construction(size, n_class, expect):
    from ganecdotes_b200.hfc_with_swav.one_shot_segmentor import OneShotSegmentor
    net = OneShotSegmentor(512, n_class, size=size)
    convs = [m for m in net.layers if isinstance(m, torch.nn.Conv2d)]
    assert [(c.in_channels, c.out_channels, c.dilation[0]) for c in convs] == expect
    assert all(c.padding[0] == c.dilation[0] for c in convs)
    assert not isinstance(net.layers[-1], torch.nn.LeakyReLU)            # no activation after the last conv
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 512, 8, 8))                             # no CPU path


def test_config_tables_match_the_reference_configs():
    """numbers of configs/segmentors/hfc_with_swav_*_config.py and configs/models/*.py"""
    from ganecdotes_b200 import configs
    c = configs.swav_config("ffhq-256")
    assert c["swav_args"]["hlen"] == 5376 and c["swav_args"]["nprototypes"] == 5000
    assert c["swav_args"]["patch_size"] == 20000 and c["swav_args"]["num_patches"] == 5
    assert c["sinkhorn_args"] == dict(source_pdf="uniform", niters=10, eps=0.005)
    assert c["swav_args"]["temperature"] == 0.01 and c["swav_args"]["trust_coeff"] == 0.01
    assert configs.seg_args("ffhq-256") == dict(size="XXS", in_ch=512)
    car = configs.swav_config("car-512")
    assert car["swav_args"]["nprototypes"] == 4000 and car["sinkhorn_args"]["eps"] == 0.01
    assert configs.seg_args("car-512")["size"] == "XS"
    assert configs.swav_config("cat-256")["sinkhorn_args"] == dict(source_pdf="image", niters=10, eps=0.003)
    pid = configs.swav_config("pidray-wrench-256")
    assert pid["swav_args"]["hlen"] == 2528 and configs.model_config("pidray-256").truncation == 0.9
    assert configs.method_for("horse-256") == "hfc_with_swav_horse"
    import pretrain
    a = pretrain.parse(["--model", "pidray-256", "--num_test_samples", "3"])
    assert a.model == "pidray-256" and a.num_test_samples == 3 and a.out_dir == "results/pretrain_default_ffhq/"


def test_config_tables_against_the_reference_config_files():
    """every per-model / per-method number of ganecdotes_b200.configs against tests/golden/reference_configs.json
    (written by tests/golden/make_config_table.py from the reference's own config files)."""
    import json
    import os
    from ganecdotes_b200 import configs
    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_configs.json")))
    assert set(table["models"]) <= set(configs.MODELS)
    for name, ref in table["models"].items():
        mc = configs.model_config(name)
        assert mc.truncation == ref["truncation"], name
        assert mc.num_latents_for_mean == ref["num_latents_for_mean"] and mc.latent_dim == ref.get("latent_dim", 512)
        assert mc.is_baggan == ref.get("is_baggan", False), name
        if name == "car-512":       # documented override: BASELINE.json asks for 512^2 features, lsun_car_512.py:8 says 256
            assert mc.image_size == 512 and ref["image_size"] == 256
        else:
            assert mc.image_size == ref["image_size"], name
        assert mc.gen_args["style_dim"] == ref["gen_args"]["style_dim"] and mc.gen_args["n_mlp"] == ref["gen_args"]["n_mlp"]
    model_of = {"hfc_with_swav_ffhq": "ffhq-256", "hfc_with_swav_cat": "cat-256", "hfc_with_swav_car": "car-512",
                "hfc_with_swav_horse": "horse-256", "hfc_with_swav_pidray": "pidray-256", "hfc_with_swav": "afhq-256"}
    assert set(table["segmentors"]) == set(model_of)
    for method, ref in table["segmentors"].items():
        model = model_of[method]
        assert configs.method_for(model) == method
        c = configs.swav_config(model)
        assert c["perturb_args"] == ref["perturb_args"], method
        assert c["perturb_args"]["n_layers"] == ref["n_hfc_layers"] == 6
        assert c["swav_args"] == ref["swav_args"], (method, {k: (c["swav_args"].get(k), v) for k, v in
                                                             ref["swav_args"].items() if c["swav_args"].get(k) != v})
        assert c["sinkhorn_args"] == ref["sinkhorn_args"], method
        assert c["layer_hf_dim"] == ref["layer_hf_dim"]
        assert configs.seg_args(model) == ref["seg_args"], method
    # the two baselines' configs (hfc_with_simclr_config.py, hfc_kmeans_config.py)
    sc, ref = configs.simclr_config(), table["baselines"]["hfc_with_simclr"]
    assert sc == ref["hfc_prep_args"] and configs.seg_args("ffhq-256", "hfc_with_simclr") == ref["seg_args"]
    kc, ref = configs.kmeans_config(), table["baselines"]["hfc_kmeans"]
    assert kc == ref["hfc_prep_args"] and configs.seg_args("ffhq-256", "hfc_kmeans") == ref["seg_args"]
    for tool in ("pliers", "hammer", "powerbank", "wrench", "handcuffs"):
        assert configs.model_config(f"pidray-{tool}-256").truncation == 0.95
        assert configs.method_for(f"pidray-{tool}-256") == "hfc_with_swav_pidray"


def test_separable_factors_of_blur_filters():
    """host side of the separable blur kernel: make_kernel filters factor exactly, a generic 4x4 filter does not"""
    from ganecdotes_b200._lib import separable_factors
    k = torch.tensor([1., 3., 3., 1.])
    k2 = k[None, :] * k[:, None]
    k2 = k2 / k2.sum() * 4
    fx, fy = separable_factors(k2)
    assert torch.equal(torch.outer(fy, fx), k2)
    a, b = torch.tensor([1., -2., 0.5, 3.]), torch.tensor([0.1, 0.2, -0.3, 0.4])
    fx, fy = separable_factors(torch.outer(a, b))
    torch.testing.assert_close(torch.outer(fy, fx), torch.outer(a, b), rtol=1e-6, atol=1e-7)
    torch.manual_seed(0)
    assert separable_factors(torch.randn(4, 4)) is None
    assert separable_factors(torch.zeros(4, 4)) is None


def test_patch_pick_rows_are_the_reference_crop():
    """sampling_method='patch': the rows of hfeat[:, :, pick:pick+P, pick:pick+P].flatten(1) (ref
    swav_clustering.py:150-158) are the pixels engine.patch_pick_rows lists, in the same order."""
    from ganecdotes_b200.hfc_with_swav.engine import patch_pick_rows
    from oracle import ganecdotes_oracle as O
    torch.manual_seed(3)
    h = w = 16
    hf = torch.randn(1, 7, h, w)
    for pick, p in [(0, 10), (3, 10), (5, 10), (12, 4), (14, 5)]:      # the last one is clipped at the border
        idx = patch_pick_rows(h, w, pick, p)
        ref = O.sample_patch_rows(hf, pick, p)
        assert torch.equal(hf[0].flatten(1)[:, idx].t(), ref)
    assert patch_pick_rows(h, w, 3, 10).numel() == 100


def test_blur_module_caches_and_refreshes_its_separable_factors():
    """Blur.separable(): factors of the registered buffer, recomputed when the buffer is overwritten (a checkpoint
    with a different filter) and None for a filter that is not an outer product (then the 2-D kernel runs)."""
    from ganecdotes_b200.stylegan2.model import Blur
    blur = Blur([1, 3, 3, 1], pad=(1, 1), upsample_factor=2)
    fx, fy = blur.separable()
    assert torch.equal(torch.outer(fy, fx), blur.kernel)
    assert blur.separable()[0] is fx                      # cached
    with torch.no_grad():
        blur.kernel.copy_(torch.outer(torch.tensor([1., 2., 2., 1.]), torch.tensor([1., 4., 4., 1.])) / 60)
    fx2, fy2 = blur.separable()
    torch.testing.assert_close(torch.outer(fy2, fx2), blur.kernel, rtol=1e-6, atol=1e-8)
    with torch.no_grad():
        blur.kernel.copy_(torch.eye(4))
    assert blur.separable() is None


def test_sinkhorn_pass_schedule_with_the_16_bit_cache(monkeypatch):
    """engine.sinkhorn_multi, host side only (the kernels are replaced by recorders): with cache16 the pass of iteration
    CACHE16_WRITE_IT writes the plane, the later ones read it, the earlier ones and every pass of a problem with
    non-uniform marginals stream the fp32 scores; the sweep direction alternates; iteration 0 is skipped when the
    score GEMM's epilogue delivered the first marginals; short chains never touch the cache."""
    calls = []

    class WS:
        partials = torch.zeros(4, 8)

        def cache16(self, chain, n):
            return ("e16", chain, n)

    def plain(s, inv_eps, first, u_in, r, c, n_total, ws, u_ll=None, reverse=False):
        calls.append(("fp32", bool(first), bool(reverse)))
        return 1

    def cached(s, inv_eps, u_in, r, c, n_total, ws, cache, write_cache, u_ll=None, reverse=False):
        assert cache[0] == "e16" and cache[2] == s.shape[0]
        calls.append(("write" if write_cache else "read", False, bool(reverse)))
        return 1

    monkeypatch.setattr(E.L, "sinkhorn_pass_parts", plain)
    monkeypatch.setattr(E.L, "sinkhorn_pass_cached_parts", cached)
    monkeypatch.setattr(E.L, "sinkhorn_reduce", lambda parts, nparts, k, out: out)
    monkeypatch.setattr(E.L, "sinkhorn_log_a", lambda u, r, **kw: u)
    s = torch.zeros(6, 8)
    w = E.CACHE16_WRITE_IT
    assert w == 2
    E.sinkhorn_multi([dict(s=s, r=None, c=None, u_first=torch.ones(8))], 10, 0.005, WS(), 6, cache16=True)
    assert [c[0] for c in calls] == ["fp32"] * (w - 1) + ["write"] + ["read"] * (10 - w - 1)
    assert [c[2] for c in calls] == [(it & 1) == 1 for it in range(1, 10)]
    assert not any(c[1] for c in calls)
    calls.clear()
    E.sinkhorn_multi([dict(s=s, r=None, c=None, u_first=None)], 10, 0.005, WS(), 6, cache16=True)
    assert calls[0] == ("fp32", True, False) and [c[0] for c in calls[1:]] == ["fp32"] * (w - 1) + ["write"] + ["read"] * 7
    calls.clear()
    E.sinkhorn_multi([dict(s=s, r=torch.ones(8) / 8, c=torch.ones(6) / 6, u_first=None)], 10, 0.005, WS(), 6, cache16=True)
    assert [c[0] for c in calls] == ["fp32"] * 10
    calls.clear()
    E.sinkhorn_multi([dict(s=s, r=None, c=None, u_first=None)], w + 1, 0.005, WS(), 6, cache16=True)
    assert [c[0] for c in calls] == ["fp32"] * (w + 1)
    calls.clear()
    E.sinkhorn_multi([dict(s=s, r=None, c=None, u_first=None)], 10, 0.005, WS(), 6)
    assert [c[0] for c in calls] == ["fp32"] * 10
