"""fused bias + leaky ReLU on sm_100a.

Mirrors `fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)` and
`FusedLeakyReLU(channel, ...)` of the reference (models/stylegan2/op/fused_act.py:11-40,
models/stylegan2/model.py:15-43, lib/gan/optim/fused_act.py:171-253) and the pybind op
`fused_bias_act` (lib/gan/optim/fused_bias_act.cpp:18-36).  Differentiable to second order like the
reference's FusedLeakyReLUFunction / ...Backward pair (lib/gan/optim/fused_act.py:27-167): the gradient
modes (`grad=1`: multiply by the slope selected by the sign of the saved output) run on the same kernel.
Bias broadcast follows models/stylegan2/op/fused_act.py:23-40 (the module this file replaces): on the LAST
dim for a 3-D input, on dim 1 otherwise.  `bias_last=False` forces dim 1 for 3-D inputs too, which is what
the BagGAN copy of the op does (lib/gan/optim/fused_act.py:171-253).
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
    # float32, float16 and float64 run natively (the reference's AT_DISPATCH_FLOATING_TYPES_AND_HALF,
    # fused_bias_act_kernel.cu:127); anything else (bfloat16) goes through float32
    native = input.dtype in (torch.float32, torch.float16, torch.float64)
    x = input if native else input.float()
    if b is not None:
        b = b.to(x.dtype).contiguous()
    if r is not None:
        r = r.to(x.dtype).contiguous()
    if x.dim() < 2 and b is not None:
        raise RuntimeError("fused_bias_act: bias needs an input with a channel dimension")
    out = L.fused_bias_act_raw(x, b, r, int(act), int(grad), float(alpha), float(scale))
    return out if native else out.to(input.dtype)


def _bias_grad(g, channel_dim_size):
    """sum over every dim but the channel dim (dim 1)"""
    dims = [0] + list(range(2, g.dim()))
    return g.sum(dim=dims) if g.dim() > 1 else g


class _FusedLReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, negative_slope, scale):
        out = fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.cfg = (negative_slope, scale, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        negative_slope, scale, has_bias = ctx.cfg
        gx, gb = _FusedLReLUGrad.apply(grad_out, out, negative_slope, scale)
        return gx, (gb if has_bias else None), None, None


class _FusedLReLUGrad(torch.autograd.Function):
    """dx = dy * scale * (out > 0 ? 1 : slope),  db = sum over non-channel dims of dx"""

    @staticmethod
    def forward(ctx, grad_out, out, negative_slope, scale):
        gx = fused_bias_act(grad_out.contiguous(), None, out, 3, 1, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.cfg = (negative_slope, scale)
        return gx, _bias_grad(gx, out.shape[1] if out.dim() > 1 else 1)

    @staticmethod
    def backward(ctx, gg_x, gg_b):
        (out,) = ctx.saved_tensors
        negative_slope, scale = ctx.cfg
        # linear in grad_out: the same map applied to (gg_x + gg_b broadcast); no dependence on `out` a.e.
        gg = fused_bias_act(gg_x.contiguous(), gg_b, out, 3, 1, negative_slope, scale)
        return gg, None, None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5, bias_last=None):
    x = input.contiguous()
    if bias_last is None:
        bias_last = x.dim() == 3                      # op/fused_act.py:26-32
    needs_grad = torch.is_grad_enabled() and (input.requires_grad or (bias is not None and bias.requires_grad))
    if bias is not None and bias_last and x.dim() == 3:
        # bias on the last dim of a 3-D input = the 2-D op on the flattened leading dims (channel dim 1)
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        y = _FusedLReLU.apply(x2, bias, negative_slope, scale) if needs_grad else \
            fused_bias_act(x2, bias, None, 3, 0, negative_slope, scale)
        return y.view(shp)
    if needs_grad:
        return _FusedLReLU.apply(x, bias, negative_slope, scale)
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
