"""CPU oracle for the ganecdotes per-pixel hidden-feature clustering path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU (torch fp32/fp64, numpy) restatement
of the reference's algorithm for the hot path named in BASELINE.json.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it, and only as the checker / CPU baseline - never as
the thing shipped.  The product path (`ganecdotes_b200`) never imports it and
fails loudly when its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors, so
this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in
the build container by `tests/golden/make_golden.py` (which imports the
unmodified reference from /root/reference) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` checks every function below against those vectors.
Third-party arithmetic on the path that is not vendored by the reference:
  * apex.parallel.LARC (version unpinned by the reference)  -> `larc_sgd_step`
    restates apex's published algorithm; parity for it is pinned only through the
    golden pretrain step, which uses the same restatement => "parity unpinned"
    for LARC itself.
  * torchvision.transforms RandomRotation/RandomHorizontalFlip (torchvision 0.26
    here) -> `rotate_flip_index_map` calls torchvision's functional ops.

All functions take explicit random draws (no hidden RNG) so that the CUDA path,
the reference and this oracle can be fed identical inputs.

Reference citations are relative to /root/reference.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SQRT2 = 2 ** 0.5


# --------------------------------------------------------------------------------------
# Generator parameters
# --------------------------------------------------------------------------------------

def stylegan2_channels(channel_multiplier: int = 2) -> Dict[int, int]:
    """models/stylegan2/model.py:484-494."""
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
            128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
            512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}


def init_generator_state(size: int, style_dim: int, n_mlp: int, seed: int,
                         channels: Optional[Dict[int, int]] = None, lr_mlp: float = 0.01,
                         randomize_small: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights in the rosinality state-dict layout
    (models/stylegan2/model.py:457-546; key names SURVEY.md §8(b)).

    The draw order is this function's own convention, so the same dict can be
    loaded into the reference Generator, this oracle and the CUDA Generator.
    `randomize_small`: also randomise the tensors the reference initialises to
    constants (biases, noise strength) so parity tests exercise them.
    """
    g = torch.Generator().manual_seed(seed)
    ch = channels or stylegan2_channels()
    sd: Dict[str, torch.Tensor] = {}

    def rn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def small(*shape, base=0.0, scale=0.1):
        if randomize_small:
            return base + scale * rn(*shape)
        return torch.full(shape, base, dtype=torch.float32)

    for i in range(n_mlp):
        sd[f"style.{i + 1}.weight"] = rn(style_dim, style_dim) / lr_mlp
        sd[f"style.{i + 1}.bias"] = small(style_dim, scale=1.0)
    sd["input.input"] = rn(1, ch[4], 4, 4)

    def styled(prefix, cin, cout, k, up):
        sd[f"{prefix}.conv.weight"] = rn(1, cout, cin, k, k)
        if up:
            kk = torch.tensor([1., 3., 3., 1.])
            kk = kk[None, :] * kk[:, None]
            sd[f"{prefix}.conv.blur.kernel"] = kk / kk.sum() * 4
        sd[f"{prefix}.conv.modulation.weight"] = rn(cin, style_dim)
        sd[f"{prefix}.conv.modulation.bias"] = small(cin, base=1.0)
        sd[f"{prefix}.noise.weight"] = small(1, scale=0.3)
        sd[f"{prefix}.activate.bias"] = small(cout)

    def torgb(prefix, cin, up):
        sd[f"{prefix}.bias"] = small(1, 3, 1, 1)
        if up:
            kk = torch.tensor([1., 3., 3., 1.])
            kk = kk[None, :] * kk[:, None]
            sd[f"{prefix}.upsample.kernel"] = kk / kk.sum() * 4
        sd[f"{prefix}.conv.weight"] = rn(1, 3, cin, 1, 1)
        sd[f"{prefix}.conv.modulation.weight"] = rn(cin, style_dim)
        sd[f"{prefix}.conv.modulation.bias"] = small(cin, base=1.0)

    styled("conv1", ch[4], ch[4], 3, False)
    torgb("to_rgb1", ch[4], False)
    log_size = int(math.log2(size))
    cin = ch[4]
    for j, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        styled(f"convs.{2 * j}", cin, cout, 3, True)
        styled(f"convs.{2 * j + 1}", cout, cout, 3, False)
        torgb(f"to_rgbs.{j}", cout, True)
        cin = cout
    num_layers = (log_size - 2) * 2 + 1
    for li in range(num_layers):
        res = (li + 5) // 2
        sd[f"noises.noise_{li}"] = rn(1, 1, 2 ** res, 2 ** res)
    return sd


def baggan_channels() -> Dict[int, int]:
    """BagGAN StyleGANGenerator channel map as read at construction time
    (models/baggan/models.py:383-390 rebinding, used at :121-122)."""
    return {4: 512, 8: 512, 16: 256, 32: 128, 64: 64, 128: 32, 256: 16, 512: 8}


def to_baggan_key(key: str) -> str:
    """rosinality Generator key -> the name the same tensor has in BagGAN's StyleGANGenerator
    (models/baggan/models.py:86-210, blocks.py:155-660)."""
    import re
    k = key
    k = re.sub(r"^style\.(\d+)\.", r"style.mapper.\1.", k)
    k = re.sub(r"^input\.input$", "const_input_block.const_block", k)
    k = re.sub(r"^conv1\.", "conv_init.", k)
    k = re.sub(r"^to_rgb1\.", "x_to_img_init.", k)
    k = re.sub(r"^convs\.(\d+)\.", r"conv_blks.\1.", k)
    k = re.sub(r"^to_rgbs\.(\d+)\.", r"x_to_img_blks.\1.", k)
    k = re.sub(r"^noises\.noise_(\d+)$", r"noise_blks.noise_\1", k)
    if k.startswith(("conv_init.", "conv_blks.")):
        k = k.replace(".conv.modulation.", ".style_block.mod.").replace(".conv.", ".style_block.")
        k = k.replace(".noise.", ".noise_block.").replace(".activate.", ".activation.")
    else:
        k = k.replace(".conv.modulation.", ".conv.mod.")
    return k


def generator_dims(sd: Dict[str, torch.Tensor]) -> Tuple[int, int, int]:
    """(size, style_dim, n_mlp) recovered from a state dict."""
    n_mlp = len([k for k in sd if k.startswith("style.") and k.endswith(".weight")])
    style_dim = sd["style.1.weight"].shape[0]
    n_up = len([k for k in sd if k.startswith("to_rgbs.") and k.count(".") == 2 and k.endswith(".bias")])
    return 4 * 2 ** n_up, style_dim, n_mlp


# --------------------------------------------------------------------------------------
# L1 ops  
# --------------------------------------------------------------------------------------

def make_fir_kernel(taps: Sequence[float]) -> torch.Tensor:
    """models/stylegan2/model.py:113-121."""
    k = torch.tensor(list(taps), dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up=1, down=1, pad=(0, 0)) -> torch.Tensor:
    """Zero-insert upsample, pad/crop, correlate with the flipped FIR, decimate.
    NCHW in/out.  Semantics of models/stylegan2/model.py:46-102 and
    lib/gan/optim/upfirdn2d_kernel.cu:114-215 (out = (in*up+p0+p1-k+down)//down)."""
    ux, uy = (up, up) if isinstance(up, int) else tuple(up)
    dx, dy = (down, down) if isinstance(down, int) else tuple(down)
    if len(pad) == 2:
        px0, px1, py0, py1 = pad[0], pad[1], pad[0], pad[1]
    else:
        px0, px1, py0, py1 = pad
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    y = x.new_zeros(n * c, 1, h * uy, w * ux)
    y[:, :, ::uy, ::ux] = x.reshape(n * c, 1, h, w)
    y = F.pad(y, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y = y[:, :, max(-py0, 0): y.shape[2] - max(-py1, 0), max(-px0, 0): y.shape[3] - max(-px1, 0)]
    y = F.conv2d(y, torch.flip(kernel, [0, 1]).to(device=x.device, dtype=x.dtype).view(1, 1, kh, kw))
    y = y[:, :, ::dy, ::dx]
    out_h = (h * uy + py0 + py1 - kh + dy) // dy
    out_w = (w * ux + px0 + px1 - kw + dx) // dx
    return y.reshape(n, c, out_h, out_w)


def fused_leaky_relu(x: torch.Tensor, bias: Optional[torch.Tensor], negative_slope: float = 0.2,
                     scale: float = SQRT2) -> torch.Tensor:
    """lrelu(x + b[c]) * scale, bias broadcast on dim 1
    (lib/gan/optim/fused_bias_act_kernel.cu:18-85, act=3 grad=0)."""
    if bias is not None:
        x = x + bias.view(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x, negative_slope) * scale


def fused_leaky_relu_op(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                        scale: float = SQRT2) -> torch.Tensor:
    """The `op/` copy of the function (models/stylegan2/op/fused_act.py:23-40): bias on the LAST dim for a
    3-D input, on dim 1 otherwise.  (That file calls input.cuda(), so it cannot run in the build container:
    restated, parity for the 3-D branch unpinned.)"""
    rest = [1] * (x.ndim - bias.ndim - 1)
    if x.ndim == 3:
        return F.leaky_relu(x + bias.view(1, *rest, bias.shape[0]), negative_slope) * scale
    return F.leaky_relu(x + bias.view(1, bias.shape[0], *rest), negative_slope) * scale


def fused_bias_act(x, bias, refer, act: int, grad: int, alpha: float, scale: float):
    """All modes of lib/gan/optim/fused_bias_act_kernel.cu:60-81 (flat tensors;
    bias indexed by dim 1)."""
    xf = x.clone()
    if bias is not None and bias.numel():
        xf = xf + bias.view(1, -1, *([1] * (x.ndim - 2)))
    ref = refer if (refer is not None and refer.numel()) else None
    mode = act * 10 + grad
    if mode == 10:
        y = xf
    elif mode == 11:
        y = xf
    elif mode == 12:
        y = torch.zeros_like(xf)
    elif mode == 30:
        y = torch.where(xf > 0, xf, xf * alpha)
    elif mode == 31:
        y = torch.where(ref > 0, xf, xf * alpha)
    elif mode == 32:
        y = torch.zeros_like(xf)
    else:
        raise ValueError("unsupported act/grad")
    return y * scale


# --------------------------------------------------------------------------------------
# Generator  
# --------------------------------------------------------------------------------------

def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """models/stylegan2/model.py:223-252."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    out = F.linear(x, weight * scale)
    if activation:
        return fused_leaky_relu(out, bias * lr_mul)
    return out + (bias * lr_mul)


def style_mlp(sd, z, lr_mlp: float = 0.01):
    """PixelNorm + n_mlp EqualLinear(fused_lrelu)  (model.py:105-110,473-482)."""
    _, _, n_mlp = generator_dims(sd)
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(n_mlp):
        x = equal_linear(x, sd[f"style.{i + 1}.weight"], sd[f"style.{i + 1}.bias"], lr_mlp, True)
    return x


def modulated_conv2d(x, style, weight, mod_w, mod_b, demodulate=True, upsample=False,
                     blur_kernel=None):
    """Per-sample modulated / demodulated conv (model.py:327-368), evaluated
    sample by sample with ordinary convolutions."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear(style, mod_w, mod_b)                      # [B, Cin]
    scale = 1 / math.sqrt(cin * k * k)
    outs = []
    for i in range(b):
        wi = scale * weight[0] * s[i].view(1, cin, 1, 1)       # [Cout, Cin, k, k]
        if demodulate:
            d = torch.rsqrt(wi.pow(2).sum([1, 2, 3]) + 1e-8)
            wi = wi * d.view(cout, 1, 1, 1)
        xi = x[i:i + 1]
        if upsample:
            o = F.conv_transpose2d(xi, wi.transpose(0, 1), padding=0, stride=2)
            kk = blur_kernel
            p = (kk.shape[0] - 2) - (k - 1)
            o = upfirdn2d(o, kk, pad=((p + 1) // 2 + 1, p // 2 + 1))
        else:
            o = F.conv2d(xi, wi, padding=k // 2)
        outs.append(o)
    return torch.cat(outs, 0)


def styled_conv(sd, prefix, x, style, noise, upsample):
    """StyledConv = ModulatedConv2d -> NoiseInjection -> FusedLeakyReLU (model.py:426-432)."""
    out = modulated_conv2d(x, style, sd[f"{prefix}.conv.weight"],
                           sd[f"{prefix}.conv.modulation.weight"],
                           sd[f"{prefix}.conv.modulation.bias"], True, upsample,
                           sd.get(f"{prefix}.conv.blur.kernel"))
    out = out + sd[f"{prefix}.noise.weight"] * noise
    return fused_leaky_relu(out, sd[f"{prefix}.activate.bias"])


def to_rgb(sd, prefix, x, style, skip):
    """ToRGB: 1x1 modulated conv without demod + bias + FIR-upsampled skip (model.py:447-454)."""
    out = modulated_conv2d(x, style, sd[f"{prefix}.conv.weight"],
                           sd[f"{prefix}.conv.modulation.weight"],
                           sd[f"{prefix}.conv.modulation.bias"], False, False)
    out = out + sd[f"{prefix}.bias"]
    if skip is not None:
        out = out + upfirdn2d(skip, sd[f"{prefix}.upsample.kernel"], up=2, pad=(2, 1))
    return out


def synthesis(sd, latent, noises: Optional[List[torch.Tensor]] = None):
    """Synthesis network on W+ latents [B, n_latent, style_dim]
    (model.py:622-648).  Returns (image, [features])."""
    size, _, _ = generator_dims(sd)
    log_size = int(math.log2(size))
    num_layers = (log_size - 2) * 2 + 1
    if noises is None:
        noises = [sd[f"noises.noise_{i}"] for i in range(num_layers)]
    b = latent.shape[0]
    out = sd["input.input"].repeat(b, 1, 1, 1)
    out = styled_conv(sd, "conv1", out, latent[:, 0], noises[0], False)
    feats = [out]
    skip = to_rgb(sd, "to_rgb1", out, latent[:, 1], None)
    i = 1
    for j in range(log_size - 2):
        out = styled_conv(sd, f"convs.{2 * j}", out, latent[:, i], noises[2 * j + 1], True)
        feats.append(out)
        out = styled_conv(sd, f"convs.{2 * j + 1}", out, latent[:, i + 1], noises[2 * j + 2], False)
        feats.append(out)
        skip = to_rgb(sd, f"to_rgbs.{j}", out, latent[:, i + 2], skip)
        i += 2
    return skip, feats


def n_latent_of(sd):
    size, _, _ = generator_dims(sd)
    return int(math.log2(size)) * 2 - 2


def generator_forward(sd, style_in, truncation=1.0, truncation_latent=None,
                      input_is_latent=False, noises=None):
    """Generator.forward for a single style entry (model.py:565-648): optional
    mapping, truncation lerp, broadcast to W+, synthesis."""
    s = style_in
    if not input_is_latent:
        s = style_mlp(sd, s)
    if truncation < 1:
        s = truncation_latent + truncation * (s - truncation_latent)
    latent = s.unsqueeze(1).repeat(1, n_latent_of(sd), 1) if s.ndim < 3 else s
    image, feats = synthesis(sd, latent, noises)
    return image, feats, latent


# --------------------------------------------------------------------------------------
# Latent-space views  
# --------------------------------------------------------------------------------------

def perturbed_wplus(sd, w, mean_latent, truncation, layer_no: int, pert_z: torch.Tensor,
                    n_layers: int, perturb_std: Sequence[float]):
    """W+ latent of one perturbed view (swav_clustering.py:593-640,
    lib/oneshot/image_augmentor.py:8-55).

    w [1,style_dim] (already mapped), pert_z [2*n_layers, style_dim] = the
    randn_like draws of image_augmentor.py:47 in order.  Returns [1, n_latent, D]
    BEFORE the second truncation applied by the feature forward."""
    wt = mean_latent + truncation * (w - mean_latent)                 # first truncation (:603-607)
    wplus = wt.unsqueeze(1).repeat(1, n_latent_of(sd), 1).clone()
    stds = [0.0] * (2 * n_layers)
    stds[2 * layer_no] = stds[2 * layer_no + 1] = perturb_std[layer_no]
    new = wplus.clone()
    for n in (2 * layer_no, 2 * layer_no + 1):
        noise_w = style_mlp(sd, pert_z[n:n + 1])
        new[:, n, :] = (1 - stds[n]) * wplus[0, n, :] + stds[n] * noise_w
    return new


def regroup_features(feats: List[torch.Tensor]) -> List[torch.Tensor]:
    """13 -> 7 per-resolution maps (image_augmentor.py:80-90, skip_const=False)."""
    n = len(feats) // 2
    return [feats[0]] + [torch.cat([feats[2 * i + 1], feats[2 * i + 2]], 1) for i in range(n)]


def pixel_feature_vectors(feats: List[torch.Tensor], hlen: int, mode: str = "nearest"):
    """Upsample every map to the largest resolution, concat on channels, slice
    (swav_clustering.py:108-130)."""
    h = max(f.shape[-2] for f in feats)
    w = max(f.shape[-1] for f in feats)
    return torch.cat([F.interpolate(f, (h, w), mode=mode) for f in feats], 1)[:, :hlen]


def view_features(sd, w, mean_latent, truncation, layer_no, pert_z, n_layers, perturb_std, hlen,
                  mode="nearest"):
    """create_hidden_features_from_perturbed_vectors (swav_clustering.py:574-656)
    with the dead first synthesis skipped."""
    wplus = perturbed_wplus(sd, w, mean_latent, truncation, layer_no, pert_z, n_layers, perturb_std)
    wplus2 = mean_latent + truncation * (wplus - mean_latent)          # second truncation (aug:75-79)
    image, feats = synthesis(sd, wplus2)
    return pixel_feature_vectors(regroup_features(feats), hlen, mode), image


# --------------------------------------------------------------------------------------
# Rotation / flip of the feature tensor  
# --------------------------------------------------------------------------------------

def rotate_flip(x: torch.Tensor, angle: float, flip: bool) -> torch.Tensor:
    """Compose([RandomRotation(10), RandomHorizontalFlip(.5)]) with the random
    draws made explicit (swav_clustering.py:98-102)."""
    import torchvision.transforms.functional as TF
    from torchvision.transforms import InterpolationMode
    y = TF.rotate(x, angle, InterpolationMode.NEAREST, False, None, [0.0] * x.shape[-3])
    return TF.hflip(y) if flip else y


def rotate_flip_index_map(h: int, w: int, angle: float, flip: bool) -> torch.Tensor:
    """int64 [h*w]: for every pixel of the transformed tensor the flat source
    pixel it copies, or -1 where the rotation fills with zero.  Obtained by
    pushing an index image through the very same torchvision ops."""
    idx = (torch.arange(h * w, dtype=torch.float32) + 1).view(1, 1, h, w)
    return rotate_flip(idx, angle, flip).round().long().flatten() - 1


# --------------------------------------------------------------------------------------
# SwAV head  
# --------------------------------------------------------------------------------------

def sample_rows(hfeat: torch.Tensor, perm: torch.Tensor, patch_size: int) -> torch.Tensor:
    """flat[:, perm][:, :patch].t()  -> [patch, D]  (swav_clustering.py:158-167,171)."""
    flat = hfeat[0].flatten(1)
    return flat[:, perm][:, :patch_size].t()


def sample_patch_rows(hfeat: torch.Tensor, pick: int, patch_size: int) -> torch.Tensor:
    """sampling_method == 'patch': hfeat[:, :, pick:pick+P, pick:pick+P] flattened row-major -> [P*P, D]
    (swav_clustering.py:150-158; the same offset `pick = np.random.choice(h - P)` on both axes, :383-385)."""
    crop = hfeat[:, :, pick:pick + patch_size, pick:pick + patch_size]
    return crop[0].flatten(1).t()


def swav_scores(rows, w_proj, w_proto, b_proto, proj_slope=None):
    """projection -> L2 normalise -> prototype (with bias)  (swav_clustering.py:171-175).
    proj_slope: projn_nw == '1-layer' (:250-256) = Linear without bias + LeakyReLU(proj_slope = 0.01)."""
    z = rows @ w_proj.t()
    if proj_slope is not None:
        z = F.leaky_relu(z, proj_slope)
    zn = F.normalize(z, p=2, dim=1)
    return zn @ w_proto.t() + b_proto


def sinkhorn_knopp(scores, niters: int, eps: float, r=None, c=None):
    """swav_clustering.py:509-544 (uniform marginals unless r / c given).
    scores [N,K] -> Q [N,K]; no max-subtraction, exactly like the reference."""
    q = torch.exp(scores / eps).t()
    q = q / torch.sum(q)
    k, n = q.shape
    r = torch.ones(k, dtype=q.dtype, device=q.device) / k if r is None else r
    c = torch.ones(n, dtype=q.dtype, device=q.device) / n if c is None else c
    for _ in range(niters):
        u = torch.sum(q, dim=1)
        q = q * (r / u).unsqueeze(1)
        q = q * (c / torch.sum(q, dim=0)).unsqueeze(0)
    return (q / torch.sum(q, dim=0, keepdim=True)).t()


def sinkhorn_log_a_cached16(scores, niters: int, eps: float, r=None, c=None, write_iter: int = 2):
    """What the CUDA path computes when the later Sinkhorn iterations run on the 16-bit cache
    (`gx_sinkhorn_pass_cached`) - a restatement of THIS repo's scheme on top of the reference's
    iteration (swav_clustering.py:509-544), so that the kernels can be checked against it to rounding while the
    distance to `sinkhorn_knopp` (the reference) is what the cache costs.  Scaling-vector form: with E = exp(S/eps),
    iteration i computes u_k = sum_n E_nk b_n, a_k = r_k / u_k, b_n = c_n / sum_k a_k E_nk.  Iteration `write_iter`
    (engine.CACHE16_WRITE_IT) stores e16_nk = half(2^15 aw_k E_nk / t_n), t_n = sum_k aw_k E_nk; later iterations use
    rho_k = a_k / aw_k, t'_n = sum_k e16_nk rho_k, u_k = (1/aw_k) sum_n e16_nk c_n / t'_n.  Returns log a [K] (float64);
    the codes are softmax_k(S/eps + log a)."""
    e = torch.exp(scores.double() / eps)
    n, k = e.shape
    r = torch.full((k,), 1.0 / k, dtype=torch.float64) if r is None else r.double()
    c = torch.full((n,), 1.0 / n, dtype=torch.float64) if c is None else c.double()
    u = e.sum(0)                                            # iteration 0 (the score GEMM's epilogue on the GPU)
    cached = niters > write_iter + 1
    for _ in range(1, niters if not cached else write_iter):
        a = r / u
        u = (e * (c / (e @ a)).unsqueeze(1)).sum(0)
    if not cached:
        return torch.log(r / u)
    aw = r / u
    p = e * aw.unsqueeze(0)
    t = p.sum(1)
    e16 = (32768.0 * p / t.unsqueeze(1)).to(torch.float16).double()
    u = (p * (c / t).unsqueeze(1)).sum(0) / aw              # iteration write_iter, exact terms
    for _ in range(write_iter + 1, niters):
        rho = (r / u) / aw
        tp = e16 @ rho
        u = (e16 * (c / tp).unsqueeze(1)).sum(0) / aw
    return torch.log(r / u)


def image_marginals(img: torch.Tensor, k: int, n: int):
    """source_pdf == 'image' marginals (swav_clustering.py:523-532)."""
    histb = torch.histc(img, n) + 1e-9
    histb[0] = histb[1]
    histb = histb / histb.sum()
    histk = torch.histc(img, k) + 1e-9
    histk[0] = histk[1]
    histk = histk / histk.sum()
    return histk, histb


def swapped_prediction_loss(p_s, p_t, q_s, q_t):
    """swav_clustering.py:547-570."""
    lst = torch.mean(torch.sum(q_s * F.log_softmax(p_t, dim=1), dim=1))
    lts = torch.mean(torch.sum(q_t * F.log_softmax(p_s, dim=1), dim=1))
    return -0.5 * (lst + lts)


def normalize_prototypes(w_proto):
    """swav_clustering.py:328-331."""
    return F.normalize(w_proto, dim=1, p=2)


def larc_sgd_step(params: List[torch.Tensor], grads: List[torch.Tensor],
                  bufs: List[Optional[torch.Tensor]], lr: float, momentum: float,
                  trust: float, weight_decay: float = 0.0, eps: float = 1e-8):
    """apex LARC(clip=False) around torch.optim.SGD(momentum)  (swav_clustering.py:286-292,458-460).
    Returns (new_params, new_bufs)."""
    new_p, new_b = [], []
    for p, g, buf in zip(params, grads, bufs):
        pn, gn = torch.norm(p), torch.norm(g)
        if pn != 0 and gn != 0:
            g = (g + weight_decay * p) * (trust * pn / (gn + pn * weight_decay + eps))
        buf = g.clone() if buf is None else momentum * buf + g
        new_p.append(p - lr * buf)
        new_b.append(buf)
    return new_p, new_b


def swav_step(rows_s: List[torch.Tensor], rows_t: List[torch.Tensor], w_proj, w_proto, b_proto,
              niters: int, eps: float, temperature: float, bufs=None, lr=0.01, momentum=0.9,
              trust=0.01, marginals=None, proj_slope=None):
    """One optimiser step of the pretrain loop (swav_clustering.py:377-460) given
    the sampled per-pixel rows of every patch: rows_s[p], rows_t[p] are [N_p, D]
    (for a batch of latents: the row-concatenation over latents = SwAV's joint /
    distributed Sinkhorn).  Prototype rows are re-normalised first.
    marginals: ((r_s, c_s), (r_t, c_t)) for source_pdf == 'image' (:523-532), else uniform.
    Returns dict(loss, grads, new params, bufs, per-patch scores/Q)."""
    w_proto = normalize_prototypes(w_proto.detach())
    wp = w_proj.detach().clone().requires_grad_(True)
    wk = w_proto.clone().requires_grad_(True)
    bk = b_proto.detach().clone().requires_grad_(True)
    loss = 0.0
    dbg = []
    for rs, rt in zip(rows_s, rows_t):
        s_s = swav_scores(rs, wp, wk, bk, proj_slope)
        s_t = swav_scores(rt, wp, wk, bk, proj_slope)
        (r_s, c_s), (r_t, c_t) = marginals if marginals is not None else ((None, None), (None, None))
        with torch.no_grad():
            q_s = sinkhorn_knopp(s_s, niters, eps, r_s, c_s)
            q_t = sinkhorn_knopp(s_t, niters, eps, r_t, c_t)
        loss = loss + swapped_prediction_loss(s_s / temperature, s_t / temperature, q_s, q_t)
        dbg.append((s_s.detach(), s_t.detach(), q_s, q_t))
    loss = loss / len(rows_s)
    loss.backward()
    grads = [wp.grad, wk.grad, bk.grad]
    new_p, new_b = larc_sgd_step([wp.detach(), wk.detach(), bk.detach()], grads,
                                 bufs or [None, None, None], lr, momentum, trust)
    return dict(loss=loss.detach(), grads=grads, params=new_p, bufs=new_b, patches=dbg,
                w_proto_normalized=w_proto)


# --------------------------------------------------------------------------------------
# SimCLR baseline head  (baseline/hfc_with_simclr/simclr_clustering.py)
# --------------------------------------------------------------------------------------

def simclr_projection(x, w1, bn_w, bn_b, w2, bn_mean=None, bn_var=None, train=True, bn_eps=1e-5):
    """Linear(no bias) -> BatchNorm1d -> LeakyReLU(0.01) -> Linear(no bias)  (simclr_clustering.py:150-160).
    train: batch statistics (biased variance), returns also the batch mean / unbiased variance that update the
    running statistics (momentum 0.1); eval: the running statistics."""
    h = x @ w1.t()
    if train:
        mean = h.mean(0)
        var_b = h.var(0, unbiased=False)
        hn = (h - mean) / torch.sqrt(var_b + bn_eps)
        stats = (mean.detach(), h.var(0, unbiased=True).detach())
    else:
        hn = (h - bn_mean) / torch.sqrt(bn_var + bn_eps)
        stats = None
    hn = hn * bn_w + bn_b
    return F.leaky_relu(hn, 0.01) @ w2.t(), stats


def simclr_loss(scores, temperature: float):
    """The reference's loops (simclr_clustering.py:235-265) in closed form.  scores [2B, C] = projection output,
    rows interleaved s_0, t_0, s_1, t_1, ...;  sim = cosine similarity / T (nn.CosineSimilarity(dim=0), eps 1e-8);
    l[i, j] = -log(exp(sim_ij) / sum_{m != i} exp(sim_im));
    loss = sum_k (l[2k-1, 2k] + l[2k, 2k-1]) / (2B).  Quirk kept: the pairs are (2k-1, 2k) - for k = 0 row -1 is
    the LAST row (Python indexing) - i.e. t_{k-1} with s_k, not the two views of one pixel."""
    n2 = scores.shape[0]
    # Second quirk kept: the reference transposes the scores to [C, 2B] (:235) and then indexes ROWS with the sample
    # counters i, j (:240) - the "sample" vectors of the similarity are the first 2B CHANNELS, each a vector over
    # the 2B samples (needs C >= 2B: 512 >= 40 in the shipped config).
    assert scores.shape[1] >= n2
    vec = scores.t()[:n2]
    nrm = vec.norm(dim=1)
    sim = (vec @ vec.t()) / torch.clamp(nrm[:, None] * nrm[None, :], min=1e-8) / temperature
    e = torch.exp(sim)
    den = e.sum(1) - e.diagonal()
    lmat = -(sim - torch.log(den)[:, None])
    k = torch.arange(n2 // 2)
    a = (2 * k - 1) % n2
    return (lmat[a, 2 * k] + lmat[2 * k, a]).sum() / n2


def simclr_step(rows_s, rows_t, params, temperature, bufs=None, bn_running=None, lr=0.01, momentum=0.9, trust=0.01):
    """One iteration of SimCLRClustering.pretrain (simclr_clustering.py:175-273) given the sampled, channel-
    normalised per-pixel rows of the two views ([B, D] each).  params = [w1, bn_w, bn_b, w2]."""
    p = [t.detach().clone().requires_grad_(True) for t in params]
    x = torch.stack([rows_s, rows_t], 1).reshape(-1, rows_s.shape[1])       # [:, ::2] = s, [:, 1::2] = t
    scores, stats = simclr_projection(x, *p)
    loss = simclr_loss(scores, temperature)
    loss.backward()
    grads = [t.grad for t in p]
    new_p, new_b = larc_sgd_step([t.detach() for t in p], grads, bufs or [None] * 4, lr, momentum, trust)
    if bn_running is None:
        bn_running = (torch.zeros_like(stats[0]), torch.ones_like(stats[1]))
    run = (0.9 * bn_running[0] + 0.1 * stats[0], 0.9 * bn_running[1] + 0.1 * stats[1])
    return dict(loss=loss.detach(), grads=grads, params=new_p, bufs=new_b, bn_running=run, scores=scores.detach())


def simclr_predict_codes(sd, w, mean_latent, truncation, params, bn_running, hlen):
    """predict_simclr_codes (simclr_clustering.py:362-401) with the projection in eval mode: channel-normalised
    per-pixel vectors -> projection -> codes [B,C,H,W], first arg-max label map."""
    wt = mean_latent + truncation * (w - mean_latent) if truncation < 1 else w
    latent = wt.unsqueeze(1).repeat(1, n_latent_of(sd), 1)
    _, feats = synthesis(sd, latent)
    hf = F.normalize(pixel_feature_vectors(feats, hlen), dim=1)
    b, d, h, ww = hf.shape
    rows = hf.permute(0, 2, 3, 1).reshape(-1, d)
    z, _ = simclr_projection(rows, *params, bn_mean=bn_running[0], bn_var=bn_running[1], train=False)
    preds = z.view(b, h, ww, -1).permute(0, 3, 1, 2)
    return preds, preds.max(1)[1]


# --------------------------------------------------------------------------------------
# Inference  
# --------------------------------------------------------------------------------------

def predict_codes(sd, w, mean_latent, truncation, w_proj, hlen, mode="nearest", proj_slope=None):
    """predict_swav_codes (swav_clustering.py:659-693): codes [B,C,H,W] fp32 and
    int64 label map [B,H,W] = first arg-max over channels."""
    wt = mean_latent + truncation * (w - mean_latent) if truncation < 1 else w
    latent = wt.unsqueeze(1).repeat(1, n_latent_of(sd), 1)
    _, feats = synthesis(sd, latent)
    hf = pixel_feature_vectors(feats, hlen, mode)
    b, d, h, ww = hf.shape
    rows = hf.permute(0, 2, 3, 1).reshape(-1, d)
    z = rows @ w_proj.t()
    if proj_slope is not None:      # projn_nw == '1-layer' (:250-256)
        z = F.leaky_relu(z, proj_slope)
    preds = z.view(b, h, ww, -1).permute(0, 3, 1, 2)
    return preds, preds.max(1)[1]


def kmeans_assign(x: torch.Tensor, centers: torch.Tensor):
    """argmin_k ||x - c_k||^2, first index on ties: the assignment step of
    sklearn KMeans.predict used by FlatKMeansHFC._layerwise_predict
    (baseline/hfc_kmeans/hfc_kmeans_clustering.py:169-208).  x [N,C], centers [K,C]
    -> int32 [N].  fp64 distances so that the oracle itself has no rounding ties."""
    d = torch.cdist(x.double(), centers.double())
    return d.argmin(1).to(torch.int32)


def kmeans_layer_maps(feats: List[torch.Tensor], centers: List[torch.Tensor], out_size: int):
    """Per-layer one-hot cluster maps resized NEAREST to out_size and mapped to
    {-1,+1} (hfc_kmeans_clustering.py:186-206).  feats = regrouped maps 1..n."""
    outs, labels = [], []
    for f, c in zip(feats, centers):
        b, ch, h, w = f.shape
        lab = kmeans_assign(f.permute(0, 2, 3, 1).reshape(-1, ch), c).view(b, 1, h, w)
        onehot = F.one_hot(lab[:, 0].long(), c.shape[0]).permute(0, 3, 1, 2).float()
        outs.append(F.interpolate(onehot, (out_size, out_size), mode="nearest") * 2 - 1)
        labels.append(lab)
    return torch.cat(outs, 1), labels


# ----------------------------------------------------------------------------------------
# one-shot segmentor head, inference forward (ref hfc_with_swav/swav_clustering.py:697-758)
# ----------------------------------------------------------------------------------------

SEGMENTOR_DILATIONS = {"XXS": [1], "XS": [1, 2, 1], "S": [1, 2, 1, 2, 1], "M": [1, 2, 4, 1, 2, 4, 1],
                       "L": [1, 2, 4, 8, 1, 2, 4, 8, 1]}                  # ref :718-724
SEGMENTOR_CHANNELS = {"XXS": [12], "XS": [16, 8], "S": [128, 64, 64, 32], "M": [128, 64, 64, 64, 64, 32],
                      "L": [128, 64, 64, 64, 64, 64, 64, 32]}            # ref :726-732


def segmentor_layers(in_ch, n_class, size):
    """[(c_in, c_out, dilation, leaky_relu_after)] exactly as the reference builds them: `zip` stops at the
    shorter list (for XXS that drops the n_class layer) and the final LeakyReLU is removed
    (ref :733-742)."""
    ch = [in_ch] + SEGMENTOR_CHANNELS[size] + [n_class]
    convs = list(zip(SEGMENTOR_DILATIONS[size], ch[:-1], ch[1:]))
    return [(ci, co, d, i + 1 < len(convs)) for i, (d, ci, co) in enumerate(convs)]


def one_shot_segmentor(state, x, n_class, size):
    """state: `layers.{2i}.weight/bias` (nn.Sequential of Conv2d, LeakyReLU pairs); x [b, in_ch, h, w]."""
    y = x
    for i, (ci, co, d, act) in enumerate(segmentor_layers(x.shape[1], n_class, size)):
        y = F.conv2d(y, state[f"layers.{2 * i}.weight"], state[f"layers.{2 * i}.bias"], padding=d, dilation=d)
        if act:
            y = F.leaky_relu(y, 0.2)
    return y
