"""upfirdn2d on sm_100a.

Mirrors `upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))` of the reference
(models/stylegan2/op/upfirdn2d.py:11-16, models/stylegan2/model.py:46-58,
lib/gan/optim/upfirdn2d.py:146-162): `up`/`down` int or (x, y), `pad` a 2-tuple applied
to both axes or a 4-tuple (x0, x1, y0, y1).  A new tensor is returned and the input is untouched.
Differentiable to any order like the reference's UpFirDn2d / UpFirDn2dBackward pair
(lib/gan/optim/upfirdn2d.py:17-143): the adjoint of an upfirdn is the upfirdn with the flipped filter,
up and down exchanged and complementary padding, so backward and double backward run on the same kernel.
Errors surface as RuntimeError like the reference's TORCH_CHECKs
(lib/gan/optim/upfirdn2d.cpp:9-15).
"""
from collections import abc

import torch

from ... import _lib as L


def _pair(v):
    if isinstance(v, abc.Iterable):
        v = tuple(v)
        if len(v) != 2:
            raise RuntimeError("upfirdn2d: up/down must be an int or a pair (x, y)")
        return int(v[0]), int(v[1])
    return int(v), int(v)


def upfirdn2d_native_layout(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """The pybind-level op of lib/gan/optim/upfirdn2d.cpp:18-39:
    input [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]."""
    if not input.is_cuda or not kernel.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")
    if not input.is_contiguous() or not kernel.is_contiguous():
        raise RuntimeError("input must be contiguous")
    if input.dim() != 4 or kernel.dim() != 2:
        raise RuntimeError("upfirdn2d: expected input [major,h,w,minor] and a 2-D kernel")
    # float32, float16 and float64 run natively (the reference's AT_DISPATCH_FLOATING_TYPES_AND_HALF,
    # upfirdn2d_kernel.cu:321), taps in the input's type; anything else (bfloat16) goes through float32
    native = input.dtype in (torch.float32, torch.float16, torch.float64)
    x = input if native else input.float()
    k = kernel if kernel.dtype == x.dtype else kernel.to(x.dtype)
    out = L.upfirdn2d_raw(x, k, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
    return out if native else out.to(input.dtype)


class _UpFirDn(torch.autograd.Function):
    """y = upfirdn(x; k, up, down, pad) on [major, h, w, 1] tensors; params = (up_x, up_y, down_x, down_y,
    pad_x0, pad_x1, pad_y0, pad_y1)."""

    @staticmethod
    def forward(ctx, x, kernel, params):
        ctx.save_for_backward(kernel)
        ctx.params = params
        ctx.in_hw = (x.shape[1], x.shape[2])
        return upfirdn2d_native_layout(x.contiguous(), kernel, *params)

    @staticmethod
    def backward(ctx, grad_out):
        (kernel,) = ctx.saved_tensors
        return _UpFirDnAdjoint.apply(grad_out, kernel, ctx.params, ctx.in_hw), None, None


class _UpFirDnAdjoint(torch.autograd.Function):
    """x_bar = upfirdn^T(y_bar): the upfirdn with the flipped filter, up <-> down, and the padding that maps
    the output grid back onto the in_h x in_w input grid."""

    @staticmethod
    def forward(ctx, grad_out, kernel, params, in_hw):
        up_x, up_y, down_x, down_y, px0, px1, py0, py1 = params
        kh, kw = kernel.shape
        in_h, in_w = in_hw
        out_h, out_w = grad_out.shape[1], grad_out.shape[2]
        adj = (down_x, down_y, up_x, up_y,
               kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
               kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        ctx.save_for_backward(kernel)
        ctx.params = params
        return upfirdn2d_native_layout(grad_out.contiguous(), torch.flip(kernel, [0, 1]).contiguous(), *adj)

    @staticmethod
    def backward(ctx, gg_in):
        (kernel,) = ctx.saved_tensors
        # the adjoint of the adjoint is the forward map
        return _UpFirDn.apply(gg_in, kernel, ctx.params), None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    up_x, up_y = _pair(up)
    down_x, down_y = _pair(down)
    pad = tuple(pad)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    if len(pad) != 4:
        raise RuntimeError("upfirdn2d: pad must have 2 or 4 entries")
    if input.dim() != 4:
        raise RuntimeError("upfirdn2d: expected a [N,C,H,W] input")
    n, c, h, w = input.shape
    x = input.contiguous().reshape(n * c, h, w, 1)
    if torch.is_grad_enabled() and input.requires_grad:
        out = _UpFirDn.apply(x, kernel.detach().contiguous(), (up_x, up_y, down_x, down_y) + tuple(int(p) for p in pad))
    else:
        out = upfirdn2d_native_layout(x, kernel.contiguous(), up_x, up_y, down_x, down_y, *pad)
    return out.view(n, c, out.shape[1], out.shape[2])
