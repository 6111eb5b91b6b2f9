"""Drop-in for the reference's `lib/gan/optim/conv2d_gradfix.py` (:1-270): `conv2d`, `conv_transpose2d` and the
`no_weight_gradients()` context, differentiable to ANY order (R1 / path-length regularisation of GAN training take
gradients of gradients), on the sm_100a kernels.

Construction: a convolution is `cols = im2col(x)` followed by a tensor-core GEMM with the flattened weight
(`gx_gemm`, 3-pass split-bf16: fp32-grade products); a transposed convolution is the GEMM followed by `col2im`.
`im2col` and `col2im` are an adjoint pair of linear maps and the GEMM is bilinear, so three autograd Functions whose
backward passes are written with each other close under differentiation: every derivative - the reference's
`Conv2dGradWeight` and its own backward included - is made of `gx_im2col`, `gx_col2im` and `gx_gemm`.

(The reference routes through cuDNN for torch 1.7 / 1.8 and falls back to `F.conv2d` on every later version,
`could_use_op`, :100-116; results are the same convolution either way.)
"""
import contextlib

import torch

from ... import _lib as L

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    """ref :12-20"""
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def ensure_tuple(xs, ndim):
    return tuple(xs) if isinstance(xs, (tuple, list)) else (xs,) * ndim


def _pad8(t):
    """zero-pad the last dim to a multiple of 8 (16-byte row pitch of the bf16 operand planes)"""
    k = t.shape[-1]
    return t if k % 8 == 0 else torch.nn.functional.pad(t, (0, 8 - k % 8))


class _MatMulNT(torch.autograd.Function):
    """a [m,k] @ b[n,k]^T on the tensor cores (gx_gemm, 3-pass split-bf16)"""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        ap, bp = _pad8(a.float().contiguous()), _pad8(b.float().contiguous())
        a_hi, a_lo = L.split_planes(ap)
        b_hi, b_lo = L.split_planes(bp)
        return L.gemm(a_hi, a_lo, b_hi, b_lo, a.shape[0], b.shape[0], ap.shape[1], 3, tag="gemm_conv2d")

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = _MatMulNT.apply(g, b.t())          # g [m,n] @ b [n,k]
        if ctx.needs_input_grad[1]:
            gb = _MatMulNT.apply(g.t(), a.t())      # g^T [n,m] @ a [m,k]
        return ga, gb


class _Im2Col(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, geom):
        kh, kw, stride, padding, dilation, ho, wo = geom
        ctx.geom, ctx.shape = geom, tuple(x.shape)
        return L.im2col(x.float().contiguous(), kh, kw, stride, padding, dilation, ho, wo, x.shape[1] * kh * kw)

    @staticmethod
    def backward(ctx, g):
        return _Col2Im.apply(g, ctx.shape, ctx.geom), None


class _Col2Im(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cols, shape, geom):
        kh, kw, stride, padding, dilation, ho, wo = geom
        ctx.geom = geom
        return L.col2im(cols.float().contiguous(), shape, kh, kw, stride, padding, dilation, ho, wo)

    @staticmethod
    def backward(ctx, g):
        return _Im2Col.apply(g, ctx.geom), None, None


def _check(input, weight):
    if not input.is_cuda or not weight.is_cuda:
        raise RuntimeError("conv2d_gradfix: CUDA tensors required (no CPU path)")
    if input.dim() != 4 or weight.dim() != 4:
        raise RuntimeError("conv2d_gradfix: expected [N,C,H,W] input and a 4-D weight")


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """ref :27-62 - same arguments and result as torch.nn.functional.conv2d"""
    _check(input, weight)
    stride, padding, dilation = ensure_tuple(stride, 2), ensure_tuple(padding, 2), ensure_tuple(dilation, 2)
    b, c, h, w = input.shape
    cout, cg, kh, kw = weight.shape
    if c != cg * groups or cout % groups:
        raise RuntimeError("conv2d_gradfix: channels do not match weight / groups")
    ho = (h + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    wo = (w + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    geom = (kh, kw, stride, padding, dilation, ho, wo)
    wt = weight.detach() if weight_gradients_disabled else weight
    og = cout // groups
    outs = []
    for gi in range(groups):
        cols = _Im2Col.apply(input[:, gi * cg:(gi + 1) * cg], geom)                        # [b*ho*wo, cg*kh*kw]
        outs.append(_MatMulNT.apply(cols, wt[gi * og:(gi + 1) * og].reshape(og, -1)))        # [b*ho*wo, og]
    y = torch.cat(outs, 1) if groups > 1 else outs[0]
    y = y.view(b, ho, wo, cout).permute(0, 3, 1, 2)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y.to(input.dtype)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """ref :66-97 - same arguments and result as torch.nn.functional.conv_transpose2d (weight [Cin, Cout/groups, kh, kw])"""
    _check(input, weight)
    stride, padding, dilation = ensure_tuple(stride, 2), ensure_tuple(padding, 2), ensure_tuple(dilation, 2)
    output_padding = ensure_tuple(output_padding, 2)
    b, cin, h, w = input.shape
    cin_w, og, kh, kw = weight.shape
    if cin != cin_w or cin % groups:
        raise RuntimeError("conv_transpose2d_gradfix: channels do not match weight / groups")
    ho = (h - 1) * stride[0] - 2 * padding[0] + dilation[0] * (kh - 1) + output_padding[0] + 1
    wo = (w - 1) * stride[1] - 2 * padding[1] + dilation[1] * (kw - 1) + output_padding[1] + 1
    # the transposed conv is the adjoint of the conv that maps [b, og, ho, wo] -> [b, cg, h, w]
    geom = (kh, kw, stride, padding, dilation, h, w)
    wt = weight.detach() if weight_gradients_disabled else weight
    cg = cin // groups
    outs = []
    for gi in range(groups):
        rows = input[:, gi * cg:(gi + 1) * cg].permute(0, 2, 3, 1).reshape(b * h * w, cg)          # [b*h*w, cg]
        wmat = wt[gi * cg:(gi + 1) * cg].reshape(cg, og * kh * kw)                                   # [cg, og*kh*kw]
        cols = _MatMulNT.apply(rows, wmat.t())                                                       # [b*h*w, og*kh*kw]
        outs.append(_Col2Im.apply(cols, (b, og, ho, wo), geom))
    y = torch.cat(outs, 1) if groups > 1 else outs[0]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y.to(input.dtype)
