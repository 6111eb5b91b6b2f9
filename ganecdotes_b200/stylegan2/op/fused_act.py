"""fused bias + leaky ReLU on sm_100a.

Mirrors `fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)` and
`FusedLeakyReLU(channel, ...)` of the reference (models/stylegan2/op/fused_act.py:11-40,
models/stylegan2/model.py:15-43, lib/gan/optim/fused_act.py:171-253) and the pybind op
`fused_bias_act` (lib/gan/optim/fused_bias_act.cpp:18-36).  Forward only.
Bias is broadcast on dim 1; for 3-D inputs the `op/` variant of the reference
broadcasts on the last dim (op/fused_act.py:26-32) - selected with `bias_last=True`.
"""
import torch
from torch import nn

from ... import _lib as L


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    """pybind-level op: empty tensors mean "absent" (fused_bias_act_kernel.cu:105-118)."""
    if not input.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")
    if not input.is_contiguous():
        raise RuntimeError("input must be contiguous")
    b = bias if (bias is not None and bias.numel()) else None
    r = refer if (refer is not None and refer.numel()) else None
    x = input if input.dtype == torch.float32 else input.float()
    if b is not None:
        b = b.to(torch.float32).contiguous()
    if r is not None:
        r = r.to(torch.float32).contiguous()
    if x.dim() < 2 and b is not None:
        raise RuntimeError("fused_bias_act: bias needs an input with a channel dimension")
    out = L.fused_bias_act_raw(x, b, r, int(act), int(grad), float(alpha), float(scale))
    return out if input.dtype == torch.float32 else out.to(input.dtype)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5, bias_last=False):
    x = input.contiguous()
    if bias is not None and bias_last and x.dim() == 3:
        # op/fused_act.py:26-32: bias on the last dim of a 3-D input
        shp = x.shape
        y = fused_bias_act(x.reshape(-1, shp[-1]), bias, None, 3, 0, negative_slope, scale)
        return y.view(shp)
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        if bias:
            self.bias = nn.Parameter(torch.zeros(channel))
        else:
            self.bias = None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
