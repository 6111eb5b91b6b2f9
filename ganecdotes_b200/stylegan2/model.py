"""StyleGAN2 generator on sm_100a - drop-in for the reference's
`models/stylegan2/model.py` Generator (rosinality layout, same parameter names so
`load_state_dict` of reference checkpoints works, same `forward` signature).

The synthesis network does not run module by module: `Generator.synthesize` drives a
fused pipeline of hand-written kernels over NHWC activations

    style/demod coefficients -> [modulated conv as tcgen05 implicit GEMM with fused
    demod+noise+bias+lrelu epilogue that also emits the next conv's modulated bf16
    operand] ; up-layers: 4 sub-pixel phase GEMMs -> fused blur+noise+bias+lrelu FIR

using the algebraic form  y = demod_o * conv(scale*W, s_i * x)  of
ModulatedConv2d.forward (ref model.py:327-368) so one weight tile serves the batch.
Features are returned as [B,C,H,W] tensors in channels_last memory (no copies).
Forward only: the clustering path runs the generator under no_grad
(ref hfc_with_swav/swav_clustering.py:593,619,671).
"""
import math
import os
import random

import torch
from torch import nn

from .. import _lib as L
from .op import FusedLeakyReLU, upfirdn2d
from .op.fused_act import fused_bias_act


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """ref model.py:32-43 - note the reference ignores `negative_slope` (hard-coded 0.2);
    reproduced."""
    return fused_bias_act(input.contiguous(), bias, None, 3, 0, 0.2, scale)


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class PixelNorm(nn.Module):
    def forward(self, input):
        return L.pixel_norm(input.contiguous().float())


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad
        self._sep = None

    def separable(self):
        """(fir_x, fir_y) of the current `kernel` buffer if it is an outer product (every filter built by
        make_kernel from a 1-D list is), else None; recomputed when the buffer changes (load_state_dict, .to())."""
        key = (self.kernel.data_ptr(), self.kernel._version, self.kernel.device)
        if self._sep is None or self._sep[0] != key:
            self._sep = (key, L.separable_factors(self.kernel))
        return self._sep[1]

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init))
        else:
            self.bias = None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        shp = input.shape
        x = input.reshape(-1, shp[-1])
        if x.dtype != torch.float32 or x.stride(1) != 1:      # strided rows (a W+ row) are read in place
            x = x.contiguous().float()
        y = L.equal_linear(x, self.weight.detach(), None if self.bias is None else self.bias.detach(), self.scale,
                           self.lr_mul, 1 if self.activation else 0)
        return y.view(*shp[:-1], y.shape[-1])

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class ModulatedConv2d(nn.Module):
    """Parameters exactly as ref model.py:272-318.  `forward(input NCHW, style)` is
    provided for op-level parity; the Generator uses the fused pipeline instead."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if downsample:
            raise NotImplementedError("downsampling convs belong to the discriminator (out of scope)")
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            pad0 = (p + 1) // 2 + factor - 1
            pad1 = p // 2 + 1
            self.blur = Blur(blur_kernel, pad=(pad0, pad1), upsample_factor=factor)
        fan_in = in_channel * kernel_size ** 2
        self.scale = 1 / math.sqrt(fan_in)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self._prep = None
        self._prep_version = None

    def prepared_small(self):
        """fp32 [9, cin, cout] weights of the direct few-channel conv (`gx_modconv_small`), per weight version"""
        ver = (self.weight._version, self.weight.data_ptr())
        if getattr(self, "_prep_small", None) is None or self._prep_small_version != ver:
            self._prep_small = L.modconv_small_weights(self.weight.detach()[0], self.scale)
            self._prep_small_version = ver
        return self._prep_small

    def prepared(self):
        """(w_hi, w_lo, wsq) planes of scale*W, cached until the weight changes."""
        ver = (self.weight._version, self.weight.data_ptr())
        if self._prep is None or self._prep_version != ver:
            self._prep = L.modconv_prepare(self.weight.detach()[0].contiguous(), self.scale)
            self._prep_version = ver
        return self._prep

    def forward(self, input, style, passes=3):
        assert self.kernel_size == 3, "standalone forward covers the 3x3 convs"
        b = input.shape[0]
        s = self.modulation(style)
        x = input.permute(0, 2, 3, 1).contiguous().float()
        x_hi, x_lo = L.modulate_split(x, s, b)
        w_hi, w_lo, wsq = self.prepared()
        demod = L.modconv_demod(wsq, s) if self.demodulate else None
        out, _, _ = L.modconv(x_hi, x_lo, w_hi, w_lo, self.out_channel, self.upsample, passes, demod=demod)
        if self.upsample:
            out, _, _ = L.blur_noise_bias_act(out, self.blur.kernel, self.blur.pad[0], self.blur.pad[1], None, None,
                                              None, 0, None, sep=self.blur.separable())
        return out.permute(0, 3, 1, 2)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))


class Generator(nn.Module):
    """ref model.py:457-648."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 channels=None):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = channels or {
            4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
            256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier,
        }
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        # numerical mode of the implicit-GEMM convs: 3 = split-bf16 (fp32-equivalent), 1 = bf16
        self.passes = 3
        self.requires_grad_(False)

    # ------------------------------------------------------------------ reference API
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    @classmethod
    def from_reference(cls, ref_generator, device="cuda"):
        """Build from an instance of the reference's Generator (or anything exposing
        size/style_dim/state_dict with the rosinality key layout)."""
        sd = ref_generator.state_dict()
        n_mlp = len([k for k in sd if k.startswith("style.") and k.endswith(".weight")])
        g = cls(ref_generator.size, ref_generator.style_dim, n_mlp)
        g.load_state_dict(sd, strict=False)
        return g.to(device)

    def forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, noise=None, randomize_noise=True):
        if not input_is_latent:
            if len(styles[0].shape) == 3:
                styles = [torch.stack([self.style(s[:, kk, :]) for kk in range(s.shape[1])], 1) for s in styles]
            else:
                styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if truncation < 1:
            tl = truncation_latent.float().contiguous()
            styles = [L.truncate(s.float().contiguous(), tl.view(-1), truncation) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            if styles[0].ndim < 3:
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            else:
                latent = styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
            latent = torch.cat([latent, latent2], 1)
        image, feats = self.synthesize(latent, noise, need_image=True)
        if return_latents:
            return image, latent
        return image, [f.permute(0, 3, 1, 2) for f in feats]

    # ------------------------------------------------------------------ fused pipeline
    # per-resolution timing tags (bench / profiling): off by default
    tag_layers = False

    def _conv_tag(self, x_hi):
        return f"modconv@{x_hi.shape[1]}" if self.tag_layers else "modconv"

    def _styled_layers(self):
        return [self.conv1] + list(self.convs)

    @torch.no_grad()
    def synthesize(self, latent, noise=None, need_image=False):
        """latent: W+ [B, n_latent, style_dim].  Returns (image NCHW or None, [NHWC fp32 features]).

        ref model.py:622-648 (block loop), :426-432 (StyledConv), :447-454 (ToRGB)."""
        latent = latent.float()
        if latent.stride(-1) != 1:          # rows of W+ are read in place (a broadcast W+ is an expand(), no copy)
            latent = latent.contiguous()
        b = latent.shape[0]
        dev = latent.device
        if noise is None:
            noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        layers = self._styled_layers()
        lat_idx = [0] + list(range(1, 1 + len(self.convs)))
        # modulation styles + demodulation coefficients for every conv (small SIMT kernels)
        styles, demods, preps = [], [], []
        for layer, li in zip(layers, lat_idx):
            s = layer.conv.modulation(latent[:, li])
            w_hi, w_lo, wsq = layer.conv.prepared()
            styles.append(s)
            demods.append(L.modconv_demod(wsq, s))
            preps.append((w_hi, w_lo))
        passes = self.passes
        want_lo = passes == 3
        # plain 3x3 layers with few channels (BagGAN 128^2 / 256^2: 32 / 16) run as a direct fp32 conv on the fp32
        # feature map of the layer before: no 64-channel padding, and that layer emits no operand planes for them
        # (measured, 16 images: 16 -> 16 at 256^2 0.60 -> 0.13 ms; 32 -> 32 at 128^2 0.16 -> 0.20 ms - one CTA of
        #  115 KB per SM - so only layers of at most 16 channels take this route)
        small = [(not ly.conv.upsample) and n > 0 and os.environ.get("GX_CONV_SMALL", "1") != "0"
                 and max(ly.conv.in_channel, ly.conv.out_channel) <= 16
                 and L.modconv_small_supported(ly.conv.in_channel, ly.conv.out_channel)
                 for n, ly in enumerate(layers)]
        ci = self.input.input
        key = (ci.data_ptr(), ci._version, ci.device)
        if getattr(self, "_const_key", None) != key:      # NHWC copy of the learned constant, refreshed when it changes
            self._const_nhwc, self._const_key = ci.detach().permute(0, 2, 3, 1).contiguous(), key
        const = self._const_nhwc
        x_hi, x_lo = L.modulate_split(const, styles[0], b, want_lo)
        feats = []
        image = None
        for n, layer in enumerate(layers):
            conv = layer.conv
            nz = noise[n]
            if nz is None:
                res = 4 * 2 ** ((n + 1) // 2)
                nz = torch.randn(b, 1, res, res, device=dev)
            nz = nz.float().contiguous()
            nxt = styles[n + 1] if n + 1 < len(layers) else None
            if n + 1 < len(layers) and small[n + 1]:
                nxt = None                      # the next layer reads this layer's fp32 map itself
            w_hi, w_lo = preps[n]
            strength = layer.noise.weight.detach()
            bias = layer.activate.bias.detach()
            if conv.upsample:
                tmp, _, _ = L.modconv(x_hi, x_lo, w_hi, w_lo, conv.out_channel, True, passes, demod=demods[n],
                                      cin_true=conv.in_channel, tag=self._conv_tag(x_hi))
                f, x_hi, x_lo = L.blur_noise_bias_act(tmp, conv.blur.kernel, conv.blur.pad[0], conv.blur.pad[1], nz,
                                                      strength, bias, 1, nxt, want_lo, sep=conv.blur.separable())
            elif small[n]:
                f, x_hi, x_lo = L.modconv_small(feats[-1], styles[n], conv.prepared_small(), demod=demods[n], noise=nz,
                                                noise_strength=strength, bias=bias, act=1, next_style=nxt,
                                                want_next_lo=want_lo, tag=f"modconv@{feats[-1].shape[1]}_small"
                                                if getattr(self, "tag_layers", False) else "modconv_small")
            else:
                f, x_hi, x_lo = L.modconv(x_hi, x_lo, w_hi, w_lo, conv.out_channel, False, passes, demod=demods[n],
                                          noise=nz, noise_strength=strength, bias=bias, act=1, next_style=nxt,
                                          want_next_lo=want_lo, cin_true=conv.in_channel, tag=self._conv_tag(x_hi))
            feats.append(f)
            if need_image and n % 2 == 0:
                rgb = self.to_rgb1 if n == 0 else self.to_rgbs[n // 2 - 1]
                s_rgb = rgb.conv.modulation(latent[:, n + 1])
                skip_up = None
                if image is not None:
                    skip_up = rgb.upsample(image).contiguous()
                image = L.torgb(f, rgb.conv.weight.detach().view(3, -1).contiguous(), rgb.conv.scale, s_rgb,
                                rgb.bias.detach().view(3).contiguous(), skip_up)
        return image, feats
