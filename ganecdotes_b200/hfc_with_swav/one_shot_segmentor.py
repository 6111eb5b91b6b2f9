"""Drop-in for the reference's `OneShotSegmentor` (hfc_with_swav/swav_clustering.py:697-758): the small
dilated 3x3 conv stack that turns the 512-channel SwAV code map into class scores, and the arg-max label
map `evaluate.py` takes from it (src/one_shot_pipeline.py:664-665).

Same constructor, same `nn.Sequential` of `Conv2d` / `LeakyReLU(0.2)` (so `state_dict` keys and pickled
checkpoints are interchangeable), including the reference's quirk that `zip` stops at the shorter list
(size 'XXS' is ONE `Conv2d(in_ch, 12, 3)` whatever `n_class` is).

The inference forward runs every layer as an implicit GEMM on the tcgen05 conv kernel (`gx_modconv` with
unmodulated planes, bias + LeakyReLU in the epilogue, next layer's operand planes emitted by the same
epilogue), on the NHWC code map `predict_swav_codes` produces - nothing goes back to the host, where the
reference runs this head (`.to('cpu')`, src/one_shot_pipeline.py:610,662).

Fine-tuning (src/one_shot_pipeline.py:540-578: CE loss + Adam through `self.segmentor(features)`): every
layer is an autograd function in the 9-tap GEMM form (forward: G = X W_all^T + stencil sum), whose backward
is two more GEMMs on the same operands: dW_all = dG^T X (MN-major, split-K) and dX = dG W_all, with dG the
9-tap spread of the output gradient (`gx_tap_spread`).  The 9-tap form stores 9 * C_out values per pixel,
which is what bounds the wide layers of sizes 'S' / 'M' / 'L' (a guard raises above 2 GB per layer).
"""
import torch
import torch.nn as nn

from .. import _lib as L

_DILATIONS = {'XXS': [1], 'XS': [1, 2, 1], 'S': [1, 2, 1, 2, 1], 'M': [1, 2, 4, 1, 2, 4, 1],
              'L': [1, 2, 4, 8, 1, 2, 4, 8, 1]}
_CHANNELS = {'XXS': [12], 'XS': [16, 8], 'S': [128, 64, 64, 32], 'M': [128, 64, 64, 64, 64, 32],
             'L': [128, 64, 64, 64, 64, 64, 64, 32]}


def _wall(weight, bias, cout_p, cin_ld):
    """[cout,cin,3,3] -> W_all [9*cout_p, cin_ld] fp32 with row (ky*3+kx)*cout_p + co, bias padded to cout_p"""
    cout, cin = weight.shape[0], weight.shape[1]
    w = weight.detach().float()
    if cout_p != cout:
        w = torch.cat([w, w.new_zeros((cout_p - cout,) + tuple(w.shape[1:]))])
    wall = w.new_zeros((9 * cout_p, cin_ld))
    wall[:, :cin] = w.permute(2, 3, 0, 1).reshape(9 * cout_p, cin)
    b = bias.detach().float()
    if cout_p != cout:
        b = torch.cat([b, b.new_zeros(cout_p - cout)])
    return wall, b.contiguous()


class _TapsConv(torch.autograd.Function):
    """3x3 (dilated) conv + bias (+ LeakyReLU 0.2) with few output channels, NHWC fp32 in / out, as one 9-tap
    GEMM + stencil sum; the backward re-uses the GEMM kernel on the saved operand planes."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, bias, dilation, act):
        b, h, w, cin = x_nhwc.shape
        cout = weight.shape[0]
        cout_p, cin_ld = (cout + 7) // 8 * 8, L.pad64(cin)
        npix = b * h * w
        x_hi, x_lo = L._planes((npix, cin_ld), x_nhwc.device, cin_ld != cin, True)
        L.split_planes(x_nhwc.detach().float().contiguous().view(npix, cin), out=(x_hi[:, :cin], x_lo[:, :cin]))
        wall, bias_p = _wall(weight, bias, cout_p, cin_ld)
        w_hi, w_lo = L.split_planes(wall, want_lo=True)
        g = L.gemm(x_hi, x_lo, w_hi, w_lo, npix, 9 * cout_p, cin_ld, 3, tag="segmentor_taps_gemm", pair=True)
        out, _, _ = L.tap_sum(g, b, h, w, cout_p, dilation, bias_p, 2 if act else 0)
        ctx.save_for_backward(x_hi, x_lo, wall, out if act else None)
        ctx.meta = (b, h, w, cin, cout, cout_p, cin_ld, int(dilation), bool(act))
        return out[..., :cout]

    @staticmethod
    def backward(ctx, grad_out):
        x_hi, x_lo, wall, out = ctx.saved_tensors
        b, h, w, cin, cout, cout_p, cin_ld, dilation, act = ctx.meta
        npix = b * h * w
        g = grad_out.float()
        if cout_p != cout:
            g = torch.cat([g, g.new_zeros((b, h, w, cout_p - cout))], dim=3)
        if act:
            g = g * torch.where(out > 0, 1.0, 0.2)
        g = g.contiguous()
        db = g.sum(dim=(0, 1, 2))[:cout]
        dg_hi, dg_lo = L.tap_spread(g, dilation)
        dwall = torch.zeros((9 * cout_p, cin_ld), dtype=torch.float32, device=g.device)
        kit = (npix + 63) // 64
        sk = max(1, min(64, kit // 8))
        L.gemm(dg_hi, dg_lo, x_hi, x_lo, 9 * cout_p, cin_ld, npix, 3, out=dwall, a_mn=True, b_mn=True, split_k=sk,
               accumulate=True, tag="segmentor_dw_gemm", pair=True)
        dweight = dwall[:, :cin].reshape(3, 3, cout_p, cin)[:, :, :cout].permute(2, 3, 0, 1).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            wt_hi, wt_lo = L.split_planes(wall, transpose=True, want_lo=True)      # [cin_ld, 9*cout_p]
            dxp = L.gemm(dg_hi, dg_lo, wt_hi, wt_lo, npix, cin_ld, 9 * cout_p, 3, tag="segmentor_dx_gemm", pair=True)
            dx = dxp[:, :cin].reshape(b, h, w, cin)
        return dx, dweight, db, None, None


class OneShotSegmentor(nn.Module):

    def __init__(self, in_ch, n_class, size='S'):
        super().__init__()
        if size == 'Lin':
            raise NotImplementedError("size='Lin' (a per-pixel Linear) is not used by any shipped config")
        if size not in _DILATIONS:
            raise AssertionError(f"size must be one of {sorted(_DILATIONS)} or 'Lin'")
        widths = [in_ch] + _CHANNELS[size] + [n_class]
        mods = []
        for d, c_in, c_out in zip(_DILATIONS[size], widths[:-1], widths[1:]):   # stops at len(dilations)
            mods += [nn.Conv2d(c_in, c_out, kernel_size=3, padding=d, dilation=d), nn.LeakyReLU(0.2, inplace=True)]
        self.layers = nn.Sequential(*mods[:-1])                                   # no activation after the last conv
        self.channels = n_class
        self.size = size
        self.passes = 3
        self._planes = {}       # conv index -> (weight version, bias version, operand planes)

    # ------------------------------------------------------------------------------------
    def _convs(self):
        return [(i, m) for i, m in enumerate(self.layers) if isinstance(m, nn.Conv2d)]

    def _prepared(self, i, conv):
        """bf16 operand planes of one conv, output channels zero-padded to a multiple of 4.
        Few output channels (9 * cout <= 256): all nine taps become columns of ONE GEMM operand
        [9 * cout, cin] and the conv is that GEMM + `gx_tap_sum` (the input is read once, not nine times);
        otherwise the implicit-GEMM conv kernel's [cout, 9 * cin] layout."""
        key = (conv.weight._version, conv.bias._version, conv.weight.data_ptr())
        hit = self._planes.get(i)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = conv.weight.detach().float()
        b = conv.bias.detach().float()
        cout, cin = w.shape[0], w.shape[1]
        pad = (-cout) % 4
        if pad:
            w = torch.cat([w, w.new_zeros((pad,) + tuple(w.shape[1:]))])
            b = torch.cat([b, b.new_zeros(pad)])
        cout_p = cout + pad
        want_lo = self.passes == 3
        if 9 * cout_p <= 256:
            cin_ld = L.pad64(cin)
            wall = w.new_zeros((9 * cout_p, cin_ld))
            wall[:, :cin] = w.permute(2, 3, 0, 1).reshape(9 * cout_p, cin)       # row = (ky*3 + kx) * cout + co
            w_hi, w_lo = L.split_planes(wall, want_lo=want_lo)
            prep = ("taps", w_hi, w_lo, b.contiguous(), cout, cout_p)
        else:
            w_hi, w_lo, _ = L.modconv_prepare(w.contiguous(), 1.0, want_lo=want_lo)
            prep = ("conv", w_hi, w_lo, b.contiguous(), cout, cout_p)
        self._planes[i] = (key, prep)
        return prep

    def _scores_nhwc(self, x, planes=None):
        L.load()
        if not x.is_cuda:
            raise RuntimeError("ganecdotes_b200.OneShotSegmentor has no CPU path (input must be a CUDA tensor)")
        b, c, h, w = x.shape
        c_ld = L.pad64(c)
        want_lo = self.passes == 3
        if planes is not None and c_ld == c:          # operand planes emitted by the producer of x
            x_hi, x_lo = planes[0].view(b, h, w, c), (planes[1].view(b, h, w, c) if want_lo else None)
        else:
            x_nhwc = x.detach().float().permute(0, 2, 3, 1).contiguous()   # free for channels_last code maps
            x_hi, x_lo = L._planes((b, h, w, c_ld), x.device, c_ld != c, want_lo)
            L.split_planes(x_nhwc.view(-1, c), out=(x_hi.view(-1, c_ld)[:, :c],
                                                   x_lo.view(-1, c_ld)[:, :c] if x_lo is not None else None))
        convs = self._convs()
        out = None
        for n, (i, conv) in enumerate(convs):
            kind, w_hi, w_lo, bias, cout, cout_p = self._prepared(i, conv)
            last = n + 1 == len(convs)
            d = conv.dilation[0]
            if kind == "taps":
                npix, k_ld = b * h * w, x_hi.shape[3]
                g = L.gemm(x_hi.view(npix, k_ld), x_lo.view(npix, k_ld) if want_lo else None, w_hi, w_lo, npix,
                           9 * cout_p, k_ld, self.passes, tag="segmentor_taps_gemm", pair=True)
                out, x_hi, x_lo = L.tap_sum(g, b, h, w, cout_p, d, bias, 0 if last else 2, want_out=last,
                                            want_planes=not last, want_lo=want_lo)
            else:
                ones = None if last else torch.ones((b, cout_p), dtype=torch.float32, device=x.device)
                out, x_hi, x_lo = L.modconv(x_hi, x_lo, w_hi, w_lo, cout_p, False, self.passes, bias=bias,
                                            act=0 if last else 2, next_style=ones, want_next_lo=want_lo,
                                            tag="segmentor_conv", cin_true=conv.in_channels, dilation=d)
        return out, cout                                                   # fp32 NHWC [b,h,w,cout_p]

    def _train_forward(self, x):
        """autograd path of the one-shot fine-tune loop"""
        L.load()
        if not x.is_cuda:
            raise RuntimeError("ganecdotes_b200.OneShotSegmentor has no CPU path (input must be a CUDA tensor)")
        y = x.float().permute(0, 2, 3, 1)
        convs = self._convs()
        for n, (i, conv) in enumerate(convs):
            g_bytes = 4 * 9 * ((conv.out_channels + 7) // 8 * 8) * y.shape[0] * y.shape[1] * y.shape[2]
            if g_bytes > 2 << 30:
                raise NotImplementedError(f"OneShotSegmentor(size={self.size!r}): the 9-tap training form of a "
                                          f"{conv.out_channels}-channel layer needs {g_bytes >> 20} MB for this "
                                          "input; fine-tune on smaller batches / crops")
            y = _TapsConv.apply(y, conv.weight, conv.bias, conv.dilation[0], n + 1 < len(convs))
        return y.permute(0, 3, 1, 2)

    def forward(self, x):
        """[b, in_ch, h, w] -> class scores [b, C_out, h, w] (channels_last memory)"""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._train_forward(x)
        out, cout = self._scores_nhwc(x)
        return out[..., :cout].permute(0, 3, 1, 2)

    @torch.no_grad()
    def predict_labels(self, x, planes=None):
        """scores + `pred.data.max(1)[1]` (src/one_shot_pipeline.py:664-665) without leaving the device:
        int64 [b, h, w], first index on ties.  `planes`: the bf16 (hi, lo) planes of x when its producer
        (`engine.predict_codes(..., want_planes=True)`) already emitted them."""
        out, cout = self._scores_nhwc(x, planes)
        b, h, w, cp = out.shape
        return L.argmax_rows(out.view(-1, cp)[:, :cout]).view(b, h, w)
