"""BagGAN-HQ generator (pidray-256 config) on the same fused synthesis pipeline.

The reference's `models/baggan/models.py::StyleGANGenerator` is the StyleGAN2 graph of
`models/stylegan2/model.py` with renamed sub-modules and a narrower channel map
(`DEFAULT_RES_TO_CHANNEL_MAP`, models.py:383-390: {4:512, 8:512, 16:256, 32:128, 64:64,
128:32, 256:16}, sum of feature channels 2528 = pidray `hlen`).  This module maps its
state-dict keys onto the drop-in `Generator` (rosinality layout) so that the pipeline's
`baggan.generator.module` (ref src/one_shot_pipeline.py:151-154) can be handed to
`SwAVClustering` unchanged.
"""
import re

from ..stylegan2.model import Generator

DEFAULT_CHL_MULTIPLIER = 2


def baggan_channels():
    """the channel map actually read by StyleGANGenerator.__init__ (models.py:383-390,121-122)"""
    m = DEFAULT_CHL_MULTIPLIER
    return {4: 512, 8: 512, 16: 256 * m // 2, 32: 128 * m // 2, 64: 64 * m // 2, 128: 32 * m // 2, 256: 16 * m // 2,
            512: 8 * m // 2, 1024: 4 * m // 2}


_RULES = [
    (r"^style\.mapper\.(\d+)\.(weight|bias)$", r"style.\1.\2"),
    (r"^const_input_block\.const_block$", r"input.input"),
    (r"^conv_init\.", r"conv1."),
    (r"^x_to_img_init\.", r"to_rgb1."),
    (r"^conv_blks\.(\d+)\.", r"convs.\1."),
    (r"^x_to_img_blks\.(\d+)\.", r"to_rgbs.\1."),
    (r"^noise_blks\.noise_(\d+)$", r"noises.noise_\1"),
]
_INNER = [
    (r"\.style_block\.mod\.", ".conv.modulation."),
    (r"\.style_block\.", ".conv."),
    (r"\.noise_block\.", ".noise."),
    (r"\.activation\.", ".activate."),
    (r"\.conv\.mod\.", ".conv.modulation."),
]


def convert_baggan_key(key):
    """BagGAN StyleGANGenerator parameter/buffer name -> rosinality Generator name (None: not
    part of the generator graph, e.g. the unused `head_m` convs, models.py:207-211)."""
    if key.startswith("head_m."):
        return None
    out = key
    for pat, rep in _RULES:
        out = re.sub(pat, rep, out)
    for pat, rep in _INNER:
        out = re.sub(pat, rep, out)
    return out


def convert_baggan_state_dict(sd):
    out = {}
    for k, v in sd.items():
        nk = convert_baggan_key(k)
        if nk is not None:
            out[nk] = v
    return out


def generator_from_baggan(ref_generator_or_state_dict, img_resolution=256, w_dim=512, mlp_layers=8, device="cuda"):
    """Build the drop-in Generator from a BagGAN StyleGANGenerator (module or state dict)."""
    sd = ref_generator_or_state_dict
    if hasattr(sd, "state_dict"):
        mod = sd
        sd = mod.state_dict()
        img_resolution = 2 ** getattr(mod, "res_log", 8)
        w_dim = getattr(mod, "w_dim", w_dim)
    sd = convert_baggan_state_dict(sd)
    n_mlp = len([k for k in sd if k.startswith("style.") and k.endswith(".weight")]) or mlp_layers
    g = Generator(img_resolution, w_dim, n_mlp, channels=baggan_channels())
    missing, unexpected = g.load_state_dict(sd, strict=False)
    unexpected = [k for k in unexpected]
    if unexpected:
        raise RuntimeError(f"unmapped BagGAN keys: {unexpected[:5]}")
    return g.to(device)
