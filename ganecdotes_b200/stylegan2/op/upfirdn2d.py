"""upfirdn2d on sm_100a.

Mirrors `upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))` of the reference
(models/stylegan2/op/upfirdn2d.py:11-16, models/stylegan2/model.py:46-58,
lib/gan/optim/upfirdn2d.py:146-162): `up`/`down` int or (x, y), `pad` a 2-tuple applied
to both axes or a 4-tuple (x0, x1, y0, y1).  Forward only (the clustering path runs
the generator under no_grad); a new tensor is returned and the input is untouched.
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
    x = input if input.dtype == torch.float32 else input.float()
    k = kernel if kernel.dtype == torch.float32 else kernel.float()
    out = L.upfirdn2d_raw(x, k, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
    return out if input.dtype == torch.float32 else out.to(input.dtype)


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
    out = upfirdn2d_native_layout(x, kernel.contiguous(), up_x, up_y, down_x, down_y, *pad)
    return out.view(n, c, out.shape[1], out.shape[2])
