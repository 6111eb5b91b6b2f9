"""One launch each of the separable blur kernel (256^2 x 128 channels, 16 image-views) and the quad upsample_sum
kernel (7 levels, 512 channels, 8 images) for `ncu --set full -k regex:"blur_sep|upsample_sum_quad"`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
fir = torch.tensor([1., 3., 3., 1.])
fir = (fir[:, None] * fir[None, :] / 64 * 4).to(dev)
sep = L.separable_factors(fir)
B, res, c = 16, 256, 128
x = torch.randn(B, res + 1, res + 1, c, device=dev)
noise = torch.randn(1, res, res, device=dev)
strength = torch.full((1,), 0.3, device=dev)
bias = torch.randn(c, device=dev)
style = torch.randn(B, c, device=dev)
for _ in range(2):
    L.blur_noise_bias_act(x, fir, 1, 1, noise, strength, bias, 1, style, sep=sep)
del x
parts = [torch.randn(8, r, r, 512, device=dev) for r in (4, 8, 16, 32, 64, 128, 256)]
out = torch.empty(8 * 256 * 256, 512, device=dev)
for _ in range(2):
    L.upsample_sum(parts, 8, 256, 256, out=out)
torch.cuda.synchronize()
print("done")
