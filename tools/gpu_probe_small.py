"""A/B timing of the small streaming kernels at the ffhq-256 hot-path shapes (16 image-views):
blur_noise_bias_act, upsample_sum (2x2-quad kernel vs the per-pixel kernel, bit-compared), split_planes.
Each figure is the median of 20 launches timed with CUDA events; the working sets exceed L2 (126 MB) for the
large layers."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
PEAK = 6529.1


def med_ms(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


fir = torch.tensor([1., 3., 3., 1.])
fir = (fir[:, None] * fir[None, :] / 64 * 4).to(dev)
B = 16
total = {"0": 0.0, "1": 0.0}
sep = L.separable_factors(fir)
for res, c in ((8, 512), (16, 512), (32, 512), (64, 512), (128, 256), (256, 128)):
    x = torch.randn(B, res + 1, res + 1, c, device=dev)
    noise = torch.randn(1, res, res, device=dev)
    strength = torch.full((1,), 0.3, device=dev)
    bias = torch.randn(c, device=dev)
    style = torch.randn(B, c, device=dev)
    nbytes = 4.0 * B * c * ((res + 1) ** 2 + res * res) + 4.0 * B * c * res * res
    outs = {}
    for mode in ("0", "1"):            # 2-D kernel / separable kernel
        os.environ["GX_BLUR_SEP"] = mode
        f = lambda: L.blur_noise_bias_act(x, fir, 1, 1, noise, strength, bias, 1, style, sep=sep)
        outs[mode] = f()[0]
        ms = med_ms(f)
        total[mode] += ms
        print(f"blur {res:4d}^2 c={c:3d} sep={mode}: {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s ({nbytes / ms / 1e6 / PEAK:.0%})")
    print("   max |sep - 2d|:", (outs["0"] - outs["1"]).abs().max().item())
    del x, outs
os.environ.pop("GX_BLUR_SEP")
print(f"blur total per 16 image-views: 2-D {total['0']:.3f} ms, separable {total['1']:.3f} ms")

# upsample_sum: 7 levels 4^2..256^2, 512 channels, 8 images (one view of the bench step); quad kernel vs per-pixel
b = 8
parts = [torch.randn(b, r, r, 512, device=dev) for r in (4, 8, 16, 32, 64, 128, 256)]
nbytes = 4.0 * 512 * (sum(p.shape[0] * p.shape[1] * p.shape[2] for p in parts) + b * 65536)
outs = {}
for quad in (0, 1):
    os.environ["GX_UPSUM_QUAD"] = str(quad)
    out = torch.empty(b * 256 * 256, 512, device=dev)
    ms = med_ms(lambda: L.upsample_sum(parts, b, 256, 256, out=out))
    print(f"upsample_sum quad={quad}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s ({nbytes / ms / 1e6 / PEAK:.0%})")
    outs[quad] = out
print("   quad == per-pixel (bit-identical):", torch.equal(outs[0], outs[1]))
assert torch.equal(outs[0], outs[1])
del outs, out, parts
# other pyramids: all levels shared (512^2 output from <= 256^2 maps), two per-pixel levels, planes output, odd batch
for sizes, oh in (((4, 8, 16, 32, 64, 128, 256), 512), ((8, 32, 64, 64), 64), ((16, 16), 16), ((4, 8, 16, 32, 64, 128, 256, 512), 512)):
    parts = [torch.randn(3, r, r, 64, device=dev) for r in sizes]
    res = {}
    for quad in (0, 1):
        os.environ["GX_UPSUM_QUAD"] = str(quad)
        hi = torch.empty(3 * oh * oh, 64, dtype=torch.bfloat16, device=dev)
        lo = torch.empty_like(hi)
        o = L.upsample_sum(parts, 3, oh, oh, planes=(hi, lo))
        res[quad] = (o, hi, lo)
    assert all(torch.equal(x, y) for x, y in zip(res[0], res[1])), (sizes, oh)
os.environ.pop("GX_UPSUM_QUAD")
print("   other pyramids (all-shared, two fine levels, 8 levels, planes): quad == per-pixel")

# split_planes: the 256^2 x 128-channel maps of 16 image-views
x = torch.randn(16 * 65536, 128, device=dev)
ms = med_ms(lambda: L.split_planes(x))
nbytes = x.numel() * 8.0
print(f"split_planes 1M x 128: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s ({nbytes / ms / 1e6 / PEAK:.0%})")
hi, lo = L.split_planes(x)
assert torch.equal(hi, x.bfloat16()) and torch.equal(lo, (x - x.bfloat16().float()).bfloat16())
print("   planes exact")
