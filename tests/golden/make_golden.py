"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (/root/reference) on CPU.  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own, so these files
are what pins the oracle (oracle/ganecdotes_oracle.py) to the reference, and -
through the oracle and directly - the CUDA path.  Every random draw the
reference makes on the path is recorded so it can be replayed.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _reference_loader as L  # noqa: E402
from oracle import ganecdotes_oracle as O  # noqa: E402

ref = L.load_reference()
torch.set_num_threads(8)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, name + ".npz"))
    print(f"wrote {name}.npz  {sz / 1024:.1f} KiB")


# ---------------------------------------------------------------------------------------
# G1: custom ops (models/stylegan2/model.py:32-102)
# ---------------------------------------------------------------------------------------
def golden_ops():
    g = torch.Generator().manual_seed(1)
    out = {}
    cases = [
        # name, shape, taps, gain, up, down, pad
        ("blur_up", (2, 5, 9, 9), [1, 3, 3, 1], 4.0, 1, 1, (1, 1)),          # Blur after transposed conv
        ("rgb_up2", (2, 3, 8, 8), [1, 3, 3, 1], 4.0, 2, 1, (2, 1)),          # Upsample of the RGB skip
        ("down2", (1, 4, 12, 10), [1, 3, 3, 1], 1.0, 1, 2, (1, 1)),          # Downsample
        ("k3", (1, 2, 7, 6), [1, 2, 1], 1.0, 1, 1, (1, 1)),
        ("crop", (1, 2, 10, 10), [1, 3, 3, 1], 1.0, 1, 1, (-1, 2)),          # negative pad crops
        ("up2down2", (1, 3, 6, 7), [1, 3, 3, 1], 1.0, 2, 2, (2, 1)),
        ("asym", (1, 2, 5, 8), [1, 4, 6, 4, 1], 1.0, (2, 1), (1, 2), (2, 1, 0, 3)),  # generic path
    ]
    for name, shape, taps, gain, up, down, pad in cases:
        x = torch.randn(*shape, generator=g)
        k = ref.model.make_kernel(taps) * gain
        if name == "asym":
            k = k + 0.01 * torch.randn(k.shape, generator=g)   # non-symmetric: flip matters
        y = ref.model.upfirdn2d(x, k, up=up, down=down, pad=pad)
        out[f"{name}_x"] = x
        out[f"{name}_k"] = k
        out[f"{name}_y"] = y
        out[f"{name}_args"] = np.array(
            list(up if isinstance(up, tuple) else (up, up)) +
            list(down if isinstance(down, tuple) else (down, down)) +
            list(pad if len(pad) == 4 else (pad[0], pad[1], pad[0], pad[1])), dtype=np.int64)
    x = torch.randn(2, 6, 5, 5, generator=g)
    b = torch.randn(6, generator=g)
    out["flr_x"], out["flr_b"] = x, b
    out["flr_y"] = ref.model.fused_leaky_relu(x, b)
    out["flr_y_nobias"] = ref.model.fused_leaky_relu(x, None)
    x2 = torch.randn(3, 6, generator=g)
    out["flr2_x"] = x2
    out["flr2_y"] = ref.model.fused_leaky_relu(x2, b)
    save("ops", **out)


# ---------------------------------------------------------------------------------------
# G2: Generator forward (models/stylegan2/model.py:565-648)
# ---------------------------------------------------------------------------------------
GEN_SIZE, GEN_STYLE, GEN_MLP, GEN_SEED = 16, 64, 2, 7


def build_reference_generator():
    sd = O.init_generator_state(GEN_SIZE, GEN_STYLE, GEN_MLP, GEN_SEED)
    gen = ref.model.Generator(GEN_SIZE, GEN_STYLE, GEN_MLP)
    missing, unexpected = gen.load_state_dict(sd, strict=True), None
    gen.eval()
    return gen, sd


def golden_generator():
    gen, sd = build_reference_generator()
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, GEN_STYLE, generator=g)
    zm = torch.randn(64, GEN_STYLE, generator=g)
    with torch.no_grad():
        w = gen.style(z)
        mean_latent = gen.style(zm).mean(0, keepdim=True)
        img, feats = gen([z], truncation=0.7, truncation_latent=mean_latent,
                         input_is_latent=False, randomize_noise=False)
        img2, latent = gen([w], return_latents=True, truncation=0.7,
                           truncation_latent=mean_latent, input_is_latent=True,
                           randomize_noise=False)
        # W+ input with per-sample noise
        wplus = torch.randn(2, gen.n_latent, GEN_STYLE, generator=g) * 0.5
        noises = [torch.randn(2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=g)
                  for i in range(gen.num_layers)]
        img3, feats3 = gen([wplus], input_is_latent=True, noise=noises)
    out = dict(z=z, zm=zm, w=w, mean_latent=mean_latent, img=img, img_latent=img2, latent=latent,
               wplus=wplus, img3=img3)
    for i, f in enumerate(feats):
        out[f"feat{i}"] = f[:, ::16]              # every 16th channel
        out[f"feat{i}_sum"] = f.double().sum(dim=(2, 3))
    for i, f in enumerate(feats3):
        out[f"feat3_{i}"] = f[:, 5::32]
    for i, n in enumerate(noises):
        out[f"noise3_{i}"] = n
    save("generator", **out)


# ---------------------------------------------------------------------------------------
# G3: SwAV pretrain steps + predict_swav_codes, RNG draws recorded
# ---------------------------------------------------------------------------------------
class Recorder:
    def __init__(self):
        self.log = []

    def wrap(self, name, fn):
        def inner(*a, **k):
            out = fn(*a, **k)
            self.log.append((name, out))
            return out
        return inner


def golden_swav(name="swav", sampling_method='random', patch=100, projn_nw='linear'):
    """sampling_method='patch' (ref swav_clustering.py:150-158, 383-385): `patch` is the side of a square crop at
    (pick, pick), pick = np.random.choice(h - patch) per patch; saved as swav_patch.npz (training part only).
    projn_nw='1-layer' (ref :250-256: Linear without bias + LeakyReLU(0.01, inplace)): saved as swav_1layer.npz."""
    gen, sd = build_reference_generator()
    swav = ref.swav
    hlen = 512 + 1024 + 1024
    nclasses, nproto, npatch, nepochs = 64, 48, 2, 2
    cfg = dict(
        perturb_args=dict(truncation=0.7, n_layers=3, n_samples=1, layer_no=None,
                          perturb_std=[1.0, 0.5, 1.0]),
        swav_args=dict(num_epochs=nepochs, num_samples=1, num_patches=npatch, sampling_method=sampling_method,
                       patch_size=patch, hf_interp='nearest', warmup_epochs=nepochs, start_warmup=0.01,
                       use_scheduler=False, base_lr=0.01, final_lr=0.0001, trust_coeff=0.01,
                       freeze_prototype_niters=313, train_args=dict(lr=0.01, momentum=0.9),
                       projn_nw=projn_nw, temperature=0.02, nprototypes=nproto, nclasses=nclasses,
                       hlen=hlen, add_local_loss=False, plot_test_images=False, epoch_print_freq=1,
                       max_masks=4),
        sinkhorn_args=dict(source_pdf='uniform', niters=10, eps=0.02),
        train=True, layer_hf_dim=[512, 1024, 1024])
    model_config = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7,
                                         latent_dim=GEN_STYLE, image_size=GEN_SIZE)
    rec = Recorder()
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    logger = types.SimpleNamespace(info=lambda *a, **k: None)

    torch.manual_seed(11)
    np.random.seed(11)
    # --- record every draw on the path -------------------------------------------------
    orig = dict(randn=torch.randn, randn_like=torch.randn_like, randperm=torch.randperm,
                choice=np.random.choice, rand=torch.rand)
    torch.randn = rec.wrap("randn", orig["randn"])
    torch.randn_like = rec.wrap("randn_like", orig["randn_like"])
    torch.randperm = rec.wrap("randperm", orig["randperm"])
    torch.rand = rec.wrap("rand", orig["rand"])
    np.random.choice = rec.wrap("choice", orig["choice"])
    from torchvision import transforms as T
    orig_gp = T.RandomRotation.get_params
    T.RandomRotation.get_params = staticmethod(rec.wrap("angle", orig_gp))
    init = {}
    orig_sgd = torch.optim.SGD

    def sgd_spy(params, **kw):
        params = list(params)
        init["params"] = [p.detach().clone() for p in params]
        return orig_sgd(params, **kw)
    torch.optim.SGD = sgd_spy
    try:
        with tempfile.TemporaryDirectory() as td:
            obj = swav.SwAVClustering(gen, model_config, logger=logger, out_dir=td, device='cpu',
                                      tb=tb, **cfg)
            mean_latent = obj.mean_latent.clone()
            n_ctor = len(rec.log)
            obj.pretrain(None, num_test_samples=0)
            n_train = len(rec.log)
            w_proj = obj.projection[0].weight.detach().clone()
            w_proto = obj.prototype.weight.detach().clone()
            b_proto = obj.prototype.bias.detach().clone()
            # inference
            wlat = gen.style(orig["randn"](1, GEN_STYLE, generator=torch.Generator().manual_seed(5)))
            preds, labels = obj.predict_swav_codes(wlat.detach())
            # direct calls of sinkhorn / loss on small random scores
            gg = torch.Generator().manual_seed(9)
            sc_s = 0.1 * orig["randn"](40, 24, generator=gg)
            sc_t = 0.1 * orig["randn"](40, 24, generator=gg)
            obj.eps = 0.005
            q_s = obj.sinkhorn_knopp(sc_s.clone(), None)
            q_t = obj.sinkhorn_knopp(sc_t.clone(), None)
            loss_direct = obj.calculate_swapped_prediction_loss(sc_s / 0.01, sc_t / 0.01, q_s, q_t)
    finally:
        torch.randn, torch.randn_like, torch.randperm = orig["randn"], orig["randn_like"], orig["randperm"]
        torch.rand = orig["rand"]
        np.random.choice = orig["choice"]
        T.RandomRotation.get_params = orig_gp
        torch.optim.SGD = orig_sgd

    # --- unpack the recorded draws per step (order: SURVEY §8 quirk 3) ---------------------
    train_log = rec.log[n_ctor:n_train]
    out = dict(mean_latent=mean_latent, mean_latent_z=rec.log[0][1],
               init_w_proj=init["params"][0], init_w_proto=init["params"][1],
               init_b_proto=init["params"][2],
               final_w_proj=w_proj, final_w_proto=w_proto, final_b_proto=b_proto,
               losses=np.array(losses), pred_w=wlat, preds=preds[:, ::4], labels=labels,
               sk_scores_s=sc_s, sk_scores_t=sc_t, sk_q_s=q_s, sk_q_t=q_t, sk_loss=loss_direct,
               cfg=np.array([hlen, nclasses, nproto, patch, npatch, nepochs, 3], dtype=np.int64),
               perturb_std=np.array([1.0, 0.5, 1.0]))
    it = iter(train_log)
    for e in range(nepochs):
        kind, z = next(it)
        assert kind == "randn" and tuple(z.shape) == (1, GEN_STYLE), (kind, z.shape)
        out[f"s{e}_z"] = z
        for v in "st":
            kind, layer = next(it)
            assert kind == "choice", kind
            out[f"s{e}_{v}_layer"] = np.int64(layer)
            pz = []
            for _ in range(6):
                kind, d = next(it)
                assert kind == "randn_like", kind
                pz.append(d)
            out[f"s{e}_{v}_pert_z"] = torch.cat(pz, 0)
        for v in "st":
            kind, ang = next(it)
            assert kind == "angle", kind
            out[f"s{e}_{v}_angle"] = np.float64(ang)
            kind, r = next(it)
            assert kind == "rand", kind
            out[f"s{e}_{v}_flip"] = np.bool_(bool(r < 0.5))
            out[f"s{e}_{v}_flip_u"] = r
        for p in range(npatch):
            kind, perm = next(it)
            if sampling_method == 'patch':
                assert kind == "choice", kind
                out[f"s{e}_pick{p}"] = np.int64(perm)
            else:
                assert kind == "randperm", kind
                out[f"s{e}_perm{p}"] = perm
    rest = list(it)
    assert not rest, [n for n, _ in rest]
    if sampling_method == 'patch':     # the training part is what differs; drop the (large) inference arrays
        # (the initial projection weights equal swav.npz's: same seed, same constructor)
        for k in ("preds", "labels", "pred_w", "sk_scores_s", "sk_scores_t", "sk_q_s", "sk_q_t", "sk_loss",
                  "init_w_proj", "mean_latent_z"):
            out.pop(k)
    if projn_nw != 'linear':           # the initial weights / mean latent equal swav.npz's (same seed, same constructor)
        for k in ("sk_scores_s", "sk_scores_t", "sk_q_s", "sk_q_t", "sk_loss", "init_w_proj", "mean_latent_z"):
            out.pop(k)
        out["preds"] = preds[:, ::8]
    save(name, **out)
    print(name, "losses", losses)


# ---------------------------------------------------------------------------------------
# G3b: SimCLR baseline head (baseline/hfc_with_simclr/simclr_clustering.py:133-281), SURVEY §8(f) rank 4
# ---------------------------------------------------------------------------------------
def golden_simclr():
    import importlib
    simclr = importlib.import_module('baseline.hfc_with_simclr.simclr_clustering')
    gen, sd = build_reference_generator()
    hlen = 512 + 1024 + 1024
    nclasses, batch, niters = 32, 6, 2
    cfg = dict(
        perturb_args=dict(truncation=0.7, n_layers=3, n_samples=1, layer_no=None, perturb_std=[1.0, 0.5, 1.0]),
        simclr_args=dict(num_iters=niters, batch_size=batch, patch_size=100, hf_interp='nearest', trust_coeff=0.01,
                         train_args=dict(lr=0.01, momentum=0.9), temperature=1.0, nclasses=nclasses, hlen=hlen,
                         epoch_print_freq=1, max_masks=4),
        train=True, layer_hf_dim=[512, 1024, 1024])
    model_config = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7,
                                         latent_dim=GEN_STYLE, image_size=GEN_SIZE)
    rec = Recorder()
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    logger = types.SimpleNamespace(info=lambda *a, **k: None)
    torch.manual_seed(13)
    np.random.seed(13)
    orig = dict(randn=torch.randn, randn_like=torch.randn_like, randperm=torch.randperm,
                choice=np.random.choice, rand=torch.rand)
    torch.randn = rec.wrap("randn", orig["randn"])
    torch.randn_like = rec.wrap("randn_like", orig["randn_like"])
    torch.randperm = rec.wrap("randperm", orig["randperm"])
    torch.rand = rec.wrap("rand", orig["rand"])
    np.random.choice = rec.wrap("choice", orig["choice"])
    from torchvision import transforms as T
    orig_gp = T.RandomRotation.get_params
    T.RandomRotation.get_params = staticmethod(rec.wrap("angle", orig_gp))
    init = {}
    orig_sgd = torch.optim.SGD

    def sgd_spy(params, **kw):
        params = list(params)
        init["params"] = [p.detach().clone() for p in params]
        return orig_sgd(params, **kw)
    torch.optim.SGD = sgd_spy
    try:
        with tempfile.TemporaryDirectory() as td:
            obj = simclr.SimCLRClustering(gen, model_config, logger=logger, out_dir=td, device='cpu', tb=tb, **cfg)
            mean_latent = obj.mean_latent.clone()
            n_ctor = len(rec.log)
            obj.pretrain(None)
            n_train = len(rec.log)
            state = {k: v.detach().clone() for k, v in obj.projection.state_dict().items()}
            wlat = gen.style(orig["randn"](1, GEN_STYLE, generator=torch.Generator().manual_seed(5)))
            obj.projection.eval()          # inference uses the running statistics
            preds, labels = obj.predict_simclr_codes(wlat.detach())
    finally:
        torch.randn, torch.randn_like, torch.randperm = orig["randn"], orig["randn_like"], orig["randperm"]
        torch.rand = orig["rand"]
        np.random.choice = orig["choice"]
        T.RandomRotation.get_params = orig_gp
        torch.optim.SGD = orig_sgd
    assert len(init["params"]) == 4, [tuple(p.shape) for p in init["params"]]
    out = dict(mean_latent=mean_latent, losses=np.array(losses),
               init_w1=init["params"][0], init_bn_w=init["params"][1], init_bn_b=init["params"][2],
               init_w2=init["params"][3],
               final_w1=state["0.weight"], final_bn_w=state["1.weight"], final_bn_b=state["1.bias"],
               final_bn_mean=state["1.running_mean"], final_bn_var=state["1.running_var"], final_w2=state["3.weight"],
               pred_w=wlat, preds=preds, labels=labels,
               cfg=np.array([hlen, nclasses, batch, niters, 3], dtype=np.int64), perturb_std=np.array([1.0, 0.5, 1.0]))
    it = iter(rec.log[n_ctor:n_train])
    for e in range(niters):
        kind, z = next(it)
        assert kind == "randn" and tuple(z.shape) == (1, GEN_STYLE), (kind, z.shape)
        out[f"s{e}_z"] = z
        for v in "st":            # per view: layer choice, 6 perturbation draws, then ITS rotation / flip (ref :187-201)
            kind, layer = next(it)
            assert kind == "choice", kind
            out[f"s{e}_{v}_layer"] = np.int64(layer)
            pz = []
            for _ in range(6):
                kind, d = next(it)
                assert kind == "randn_like", kind
                pz.append(d)
            out[f"s{e}_{v}_pert_z"] = torch.cat(pz, 0)
            kind, ang = next(it)
            assert kind == "angle", kind
            out[f"s{e}_{v}_angle"] = np.float64(ang)
            kind, r = next(it)
            assert kind == "rand", kind
            out[f"s{e}_{v}_flip"] = np.bool_(bool(r < 0.5))
        kind, perm = next(it)
        assert kind == "randperm", kind
        out[f"s{e}_perm"] = perm
    rest = list(it)
    assert not rest, [n for n, _ in rest]
    save("simclr", **out)
    print("simclr losses", losses)


# ---------------------------------------------------------------------------------------
# G3c: full-size ffhq-256 label map (BASELINE config 1 geometry): predict_swav_codes of the unmodified reference,
# Generator(256, 512, 8) with seeded random-init weights, hlen 5376, 512 code channels.  The 11 MB projection
# matrix is regenerated from its seed by the tests (full_size_projection below), only the label map, the
# top-2 margins (to qualify near-ties) and a strided sample of the codes are stored.
# ---------------------------------------------------------------------------------------
FULL_GEN_SEED, FULL_PROJ_SEED, FULL_W_SEED = 42, 1234, 77


def full_size_projection():
    g = torch.Generator().manual_seed(FULL_PROJ_SEED)
    return torch.randn(512, 5376, generator=g) / 5376 ** 0.5


def golden_labelmap_full(size=256, name="labelmap_ffhq256", nimg=2):
    """size=512: BASELINE config 3 geometry (car-512): a 512^2 generator whose features sum to 5504 channels, sliced
    to hlen = 5376 AFTER upsampling to 512^2 (the two 512^2-native 64-channel maps lose all but
    their first channels, everything is sampled at 512^2)."""
    sd = O.init_generator_state(size, 512, 8, FULL_GEN_SEED)
    gen = ref.model.Generator(size, 512, 8)
    gen.load_state_dict(sd, strict=True)
    gen.eval()
    swav = ref.swav
    cfg = dict(perturb_args=dict(truncation=0.7, n_layers=6, n_samples=1, layer_no=None, perturb_std=[1.0] * 6),
               swav_args=dict(hf_interp='nearest', hlen=5376, nclasses=512, nprototypes=5000,
                              sampling_method='random', patch_size=20000, num_patches=5, projn_nw='linear',
                              temperature=0.01),
               sinkhorn_args=dict(source_pdf='uniform', niters=10, eps=0.005),
               train=False, layer_hf_dim=[512, 1024, 1024, 1024, 1024, 512, 256])
    model_config = types.SimpleNamespace(num_latents_for_mean=256, truncation=0.7, latent_dim=512, image_size=size)
    logger = types.SimpleNamespace(info=lambda *a, **k: None)
    g = torch.Generator().manual_seed(FULL_W_SEED)
    zm = torch.randn(256, 512, generator=g)
    z = torch.randn(2, 512, generator=g)[:nimg]
    orig_randn = torch.randn
    torch.randn = lambda *a, **k: zm.clone() if a[:2] == (256, 512) else orig_randn(*a, **k)   # mean_latent draws
    try:
        with tempfile.TemporaryDirectory() as td:
            obj = swav.SwAVClustering(gen, model_config, logger=logger, out_dir=td, device='cpu', tb=None, **cfg)
    finally:
        torch.randn = orig_randn
    obj.projection = torch.nn.Sequential(torch.nn.Linear(5376, 512, bias=False))
    with torch.no_grad():
        obj.projection[0].weight.copy_(full_size_projection())
    # zm and z are the first two draws of torch.Generator().manual_seed(FULL_W_SEED): the tests regenerate them
    out = dict(mean_latent=obj.mean_latent,
               seeds=np.array([FULL_GEN_SEED, FULL_PROJ_SEED, FULL_W_SEED], dtype=np.int64))
    with torch.no_grad():
        wl = gen.style(z)
        for i in range(z.shape[0]):
            preds, labels = obj.predict_swav_codes(wl[i:i + 1])
            top2 = preds.topk(2, dim=1).values
            out[f"labels{i}"] = labels.to(torch.int16)
            margin = top2[:, 0] - top2[:, 1]
            if size == 256:
                out[f"margin{i}"] = margin.to(torch.float16)
            else:                       # 512^2: only whether a pixel is a near-tie (packed bits)
                out[f"neartie{i}"] = np.packbits((margin < 1e-3 * preds.abs().max()).numpy())
            out[f"absmax{i}"] = preds.abs().max()
            out[f"preds{i}_sample"] = preds[:, ::16, ::8, ::8] if size == 256 else preds[:, ::16, ::16, ::16]
    save(name, **out)


# ---------------------------------------------------------------------------------------
# G3d: BASELINE config 1 at full size - ONE pretrain step of the unmodified reference on CPU: ffhq-256 geometry
# (Generator(256, 512, 8), hlen 5376, 512 classes, 5000 prototypes, 5 patches x 20000 px, eps 0.005, T 0.01),
# random-init weights from seeds.  The 21 MB of head weights are regenerated from their seeds by the tests; the
# file holds the recorded draws (perms truncated to the 20000 picks that are used), the loss, strided samples of
# the updated weights and the update norms.
# ---------------------------------------------------------------------------------------
def full_size_head():
    g = torch.Generator().manual_seed(FULL_PROJ_SEED)
    w_proj = torch.randn(512, 5376, generator=g) / 5376 ** 0.5
    w_proto = torch.randn(5000, 512, generator=g) / 512 ** 0.5
    b_proto = 0.01 * torch.randn(5000, generator=g)
    return w_proj, w_proto, b_proto


def golden_pretrain_full():
    sd = O.init_generator_state(256, 512, 8, FULL_GEN_SEED)
    gen = ref.model.Generator(256, 512, 8)
    gen.load_state_dict(sd, strict=True)
    gen.eval()
    swav = ref.swav
    cfg = dict(perturb_args=dict(truncation=0.7, n_layers=6, n_samples=1, layer_no=None, perturb_std=[1.0] * 6),
               swav_args=dict(num_epochs=1, num_samples=1, num_patches=5, sampling_method='random', patch_size=20000,
                              hf_interp='nearest', warmup_epochs=1, start_warmup=0.01, use_scheduler=False,
                              base_lr=0.01, final_lr=0.0001, trust_coeff=0.01, freeze_prototype_niters=313,
                              train_args=dict(lr=0.01, momentum=0.9), projn_nw='linear', temperature=0.01,
                              nprototypes=5000, nclasses=512, hlen=5376, add_local_loss=False,
                              plot_test_images=False, epoch_print_freq=1, max_masks=4),
               sinkhorn_args=dict(source_pdf='uniform', niters=10, eps=0.005),
               train=True, layer_hf_dim=[512, 1024, 1024, 1024, 1024, 512, 256])
    model_config = types.SimpleNamespace(num_latents_for_mean=256, truncation=0.7, latent_dim=512, image_size=256)
    init = full_size_head()          # before the recorder wraps torch.randn
    rec = Recorder()
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    logger = types.SimpleNamespace(info=lambda *a, **k: None)
    torch.manual_seed(21)
    np.random.seed(21)
    orig = dict(randn=torch.randn, randn_like=torch.randn_like, randperm=torch.randperm,
                choice=np.random.choice, rand=torch.rand)
    torch.randn = rec.wrap("randn", orig["randn"])
    torch.randn_like = rec.wrap("randn_like", orig["randn_like"])
    torch.randperm = rec.wrap("randperm", orig["randperm"])
    torch.rand = rec.wrap("rand", orig["rand"])
    np.random.choice = rec.wrap("choice", orig["choice"])
    from torchvision import transforms as T
    orig_gp = T.RandomRotation.get_params
    T.RandomRotation.get_params = staticmethod(rec.wrap("angle", orig_gp))
    orig_sgd = torch.optim.SGD

    def sgd_spy(params, **kw):      # the head starts from the seeded weights (as if loaded from a checkpoint)
        params = list(params)
        with torch.no_grad():
            for p, v in zip(params, init):
                p.copy_(v)
        return orig_sgd(params, **kw)
    torch.optim.SGD = sgd_spy
    try:
        with tempfile.TemporaryDirectory() as td:
            obj = swav.SwAVClustering(gen, model_config, logger=logger, out_dir=td, device='cpu', tb=tb, **cfg)
            mean_latent = obj.mean_latent.clone()
            n_ctor = len(rec.log)
            obj.pretrain(None, num_test_samples=0)
            n_train = len(rec.log)
            fin = [obj.projection[0].weight.detach().clone(), obj.prototype.weight.detach().clone(),
                   obj.prototype.bias.detach().clone()]
    finally:
        torch.randn, torch.randn_like, torch.randperm = orig["randn"], orig["randn_like"], orig["randperm"]
        torch.rand = orig["rand"]
        np.random.choice = orig["choice"]
        T.RandomRotation.get_params = orig_gp
        torch.optim.SGD = orig_sgd
    ref_init = [init[0], torch.nn.functional.normalize(init[1], dim=1), init[2]]
    # the mean-latent draws are torch.randn(256, 512) right after torch.manual_seed(21): the tests regenerate them
    assert tuple(rec.log[0][1].shape) == (256, 512)
    out = dict(mean_latent=mean_latent, losses=np.array(losses),
               seeds=np.array([FULL_GEN_SEED, FULL_PROJ_SEED, 21], dtype=np.int64),
               final_w_proj_sample=fin[0][::16, ::64], final_w_proto_sample=fin[1][::50, ::16],
               final_b_proto_sample=fin[2][::10],
               update_norms=np.array([(f - i).norm().item() for f, i in zip(fin, ref_init)]),
               delta_w_proj_sample=(fin[0] - ref_init[0])[::16, ::64],
               delta_w_proto_sample=(fin[1] - ref_init[1])[::50, ::16])
    it = iter(rec.log[n_ctor:n_train])
    kind, z = next(it)
    assert kind == "randn" and tuple(z.shape) == (1, 512), (kind, z.shape)
    out["z"] = z
    for v in "st":
        kind, layer = next(it)
        assert kind == "choice", kind
        out[f"{v}_layer"] = np.int64(layer)
        pz = []
        for _ in range(12):
            kind, d = next(it)
            assert kind == "randn_like", kind
            pz.append(d)
        out[f"{v}_pert_z"] = torch.cat(pz, 0)
    for v in "st":
        kind, ang = next(it)
        assert kind == "angle", kind
        out[f"{v}_angle"] = np.float64(ang)
        kind, r = next(it)
        assert kind == "rand", kind
        out[f"{v}_flip"] = np.bool_(bool(r < 0.5))
    for p in range(5):
        kind, perm = next(it)
        assert kind == "randperm", kind
        out[f"perm{p}"] = perm[:20000].to(torch.int32)
    rest = list(it)
    assert not rest, [n for n, _ in rest]
    save("pretrain_ffhq256", **out)
    print("full-size losses", losses, "update norms", out["update_norms"])


# ---------------------------------------------------------------------------------------
# G4: BagGAN generator (models/baggan/models.py:86-379), pidray channel map
# ---------------------------------------------------------------------------------------
def golden_baggan():
    """The real `lib/gan/optim` python modules are imported with `cpp_extension.load` neutralised
    (no CUDA JIT in this container): on CPU tensors they take their own native branches
    (upfirdn2d.py:156-157, fused_act.py:234-248), which is the path the reference's CPU run uses."""
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: types.SimpleNamespace()
    import models.baggan.models as bm
    sys.path.insert(0, ROOT)
    from oracle import ganecdotes_oracle as O2
    size = 256
    ch = O2.baggan_channels()
    sd = O2.init_generator_state(size, 512, 8, 13, channels=ch)
    gen = bm.StyleGANGenerator((512, 512), size)
    ref_sd = {O2.to_baggan_key(k): v for k, v in sd.items()}
    missing, unexpected = gen.load_state_dict(ref_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("head_m.") for k in missing), missing
    gen.eval()
    g = torch.Generator().manual_seed(4)
    w = torch.randn(2, 512, generator=g) * 0.7
    zm = torch.randn(32, 512, generator=g)
    with torch.no_grad():
        mean_latent = gen.style(zm).mean(0, keepdim=True)
        wz = gen.style(torch.randn(2, 512, generator=g))
        img, feats = gen([wz], truncation=0.9, truncation_latent=mean_latent, input_is_latent=True,
                         randomize_noise=False)
    out = dict(w=wz, zm=zm, mean_latent=mean_latent, img=img, size=np.int64(size),
               chans=np.array([f.shape[1] for f in feats], dtype=np.int64))
    for i, f in enumerate(feats):
        cs = 16 if f.shape[1] >= 128 else 4
        sp = 1 if f.shape[-1] <= 16 else (2 if f.shape[-1] <= 32 else f.shape[-1] // 16)
        out[f"feat{i}"] = f[:, 1::cs, ::sp, ::sp]
        out[f"feat{i}_sum"] = f.double().sum(dim=(2, 3))
    out["img"] = img[:, :, ::4, ::4]
    save("baggan", **out)


# ---------------------------------------------------------------------------------------
# G5: one-shot segmentor head (hfc_with_swav/swav_clustering.py:697-758), inference forward
# ---------------------------------------------------------------------------------------

def golden_segmentor():
    """Weights and input are seeded (default nn.Conv2d init in construction order), so only the outputs and
    a checksum of the parameters are stored: the tests rebuild both from the seeds."""
    out = {}
    x = torch.randn(2, 512, 24, 24, generator=torch.Generator().manual_seed(21))
    for size, n_class in (("XXS", 5), ("XS", 6), ("S", 7)):
        torch.manual_seed(100 + n_class)
        net = ref.swav.OneShotSegmentor(512, n_class, size=size).eval()
        sd = net.state_dict()
        out[f"{size}_keys"] = np.array(list(sd.keys()))
        out[f"{size}_shapes"] = np.array([";".join(str(d) for d in v.shape) for v in sd.values()])
        out[f"{size}_param_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
        with torch.no_grad():
            y = net(x)
        out[f"{size}_y"] = y
        out[f"{size}_labels"] = y.max(1)[1]
        out[f"{size}_nclass"] = n_class
    save("segmentor", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "segmentor":
        golden_segmentor()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pretrain_full":
        golden_pretrain_full()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "labelmap_full":
        golden_labelmap_full()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "labelmap_car512":
        golden_labelmap_full(512, "labelmap_car512", 1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "simclr":
        golden_simclr()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "swav_patch":
        golden_swav("swav_patch", 'patch', 10)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "swav_1layer":
        golden_swav("swav_1layer", 'random', 100, '1-layer')
        sys.exit(0)
    golden_ops()
    golden_generator()
    golden_swav()
    golden_swav("swav_patch", 'patch', 10)
    golden_swav("swav_1layer", 'random', 100, '1-layer')
    golden_simclr()
    golden_labelmap_full()
    golden_labelmap_full(512, "labelmap_car512", 1)
    golden_pretrain_full()
    golden_baggan()
    golden_segmentor()
