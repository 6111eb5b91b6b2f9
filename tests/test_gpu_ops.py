"""GPU parity tests of the individual kernels, called through the C ABI (ctypes) and
checked against the golden vectors of the reference and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ganecdotes_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


@pytest.fixture(scope="module")
def L():
    from ganecdotes_b200 import _lib
    _lib.load()
    return _lib


def planes(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


# ------------------------------------------------------------------------------- L0 ops
@pytest.mark.parametrize("name", ["blur_up", "rgb_up2", "down2", "k3", "crop", "up2down2", "asym"])
def test_upfirdn2d_golden(L, name):
    from ganecdotes_b200.stylegan2.op import upfirdn2d
    g = load("ops")
    a = [int(v) for v in g[f"{name}_args"]]
    y = upfirdn2d(g[f"{name}_x"].cuda(), g[f"{name}_k"].cuda(), up=(a[0], a[1]), down=(a[2], a[3]), pad=tuple(a[4:8]))
    assert y.shape == g[f"{name}_y"].shape
    # fp32, same taps, different summation order / FMA contraction than the reference conv2d
    torch.testing.assert_close(y.cpu(), g[f"{name}_y"], rtol=1e-5, atol=1e-6)


def test_upfirdn2d_native_layout_minor(L):
    """the pybind-level layout [major,h,w,minor] with minor > 1 (ref upfirdn2d.cpp:18-39)"""
    from ganecdotes_b200.stylegan2.op import upfirdn2d_native_layout
    torch.manual_seed(0)
    x = torch.randn(3, 9, 7, 5)
    k = O.make_fir_kernel([1, 3, 3, 1]) * 4
    ref = O.upfirdn2d(x.permute(0, 3, 1, 2).contiguous(), k, up=2, down=1, pad=(2, 1)).permute(0, 2, 3, 1)
    y = upfirdn2d_native_layout(x.cuda(), k.cuda(), 2, 2, 1, 1, 2, 1, 2, 1)
    torch.testing.assert_close(y.cpu(), ref.contiguous(), rtol=1e-5, atol=1e-6)


def test_upfirdn2d_errors(L):
    from ganecdotes_b200.stylegan2.op import upfirdn2d_native_layout
    k = O.make_fir_kernel([1, 3, 3, 1])
    with pytest.raises(RuntimeError):
        upfirdn2d_native_layout(torch.randn(1, 4, 4, 1), k, 1, 1, 1, 1, 0, 0, 0, 0)       # CPU tensor
    with pytest.raises(RuntimeError):
        upfirdn2d_native_layout(torch.randn(1, 4, 4, 2).cuda()[..., :1], k.cuda(), 1, 1, 1, 1, 0, 0, 0, 0)


def test_fused_leaky_relu_golden(L):
    from ganecdotes_b200.stylegan2.op import fused_leaky_relu, fused_bias_act, FusedLeakyReLU
    g = load("ops")
    y = fused_leaky_relu(g["flr_x"].cuda(), g["flr_b"].cuda())
    torch.testing.assert_close(y.cpu(), g["flr_y"], rtol=1e-6, atol=1e-7)
    y = fused_leaky_relu(g["flr_x"].cuda(), None)
    torch.testing.assert_close(y.cpu(), g["flr_y_nobias"], rtol=1e-6, atol=1e-7)
    y = fused_leaky_relu(g["flr2_x"].cuda(), g["flr_b"].cuda())
    torch.testing.assert_close(y.cpu(), g["flr2_y"], rtol=1e-6, atol=1e-7)
    # 3-D input: the op/ module this file replaces broadcasts the bias on the LAST dim by default
    # (models/stylegan2/op/fused_act.py:23-40); bias_last=False is the lib/gan/optim (dim 1) behaviour
    from oracle import ganecdotes_oracle as O
    x3 = torch.randn(4, 6, 6, generator=torch.Generator().manual_seed(3))
    b6 = g["flr_b"]
    torch.testing.assert_close(fused_leaky_relu(x3.cuda(), b6.cuda()).cpu(), O.fused_leaky_relu_op(x3, b6),
                               rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(fused_leaky_relu(x3.cuda(), b6.cuda(), bias_last=False).cpu(),
                               O.fused_leaky_relu(x3, b6), rtol=1e-6, atol=1e-7)
    xg, bg = x3.cuda().requires_grad_(True), b6.cuda().requires_grad_(True)
    xr, br = x3.clone().requires_grad_(True), b6.clone().requires_grad_(True)
    go = torch.randn(4, 6, 6, generator=torch.Generator().manual_seed(4))
    fused_leaky_relu(xg, bg).backward(go.cuda())
    O.fused_leaky_relu_op(xr, br).backward(go)
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(bg.grad.cpu(), br.grad, rtol=1e-5, atol=1e-6)
    m = FusedLeakyReLU(6).cuda()
    assert "bias" in dict(m.named_parameters())
    with torch.no_grad():
        m.bias.copy_(g["flr_b"])
    torch.testing.assert_close(m(g["flr_x"].cuda()).cpu(), g["flr_y"], rtol=1e-6, atol=1e-7)
    # all modes of the native op, vectorised and scalar paths
    torch.manual_seed(1)
    for shape in [(2, 6, 8, 8), (3, 5, 3)]:
        x = torch.randn(*shape)
        b = torch.randn(shape[1])
        r = torch.randn(*shape)
        for act, grad in [(1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (3, 2)]:
            ref = O.fused_bias_act(x, b, r, act, grad, 0.2, 1.5)
            out = fused_bias_act(x.cuda(), b.cuda(), r.cuda(), act, grad, 0.2, 1.5)
            torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-7)
    assert fused_bias_act(torch.empty(0, 4).cuda(), torch.empty(0).cuda(), torch.empty(0).cuda(), 3, 0, 0.2,
                          1.0).numel() == 0


# ------------------------------------------------------------------------------- GEMM engine
@pytest.mark.parametrize("m,n,k,passes,a_mn,b_mn,split,bias", [
    (128, 128, 64, 1, 0, 0, 1, 0),
    (300, 72, 200, 1, 0, 0, 1, 1),          # ragged everywhere
    (1000, 520, 5376, 3, 0, 0, 1, 0),       # projection-shaped
    (777, 5000, 512, 3, 0, 0, 1, 1),        # prototype-shaped, N not a tile multiple
    (256, 512, 5000, 1, 0, 0, 1, 0),        # dZn-shaped, K not a multiple of 64
    (5000, 512, 3000, 1, 1, 1, 7, 0),       # gWk-shaped, both MN-major, split-K
    (520, 5376, 4096, 3, 1, 1, 4, 0),       # gWp-shaped
    (1, 8, 8, 3, 0, 0, 1, 1),               # minimum sizes
    (4000, 512, 1024, 1, 0, 0, 1, 1),       # 256-row CTA tiles, K-major, ragged M
    (304, 512, 2048, 1, 1, 0, 3, 0),        # 256-row tiles, MN-major A only (MN-major rows need a 16-byte pitch)
])
def test_gemm_vs_fp64(L, m, n, k, passes, a_mn, b_mn, split, bias):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(n, k, device="cuda")
    ah, al = planes(a.t().contiguous() if a_mn else a)
    bh, bl = planes(b.t().contiguous() if b_mn else b)
    bv = torch.randn(n, device="cuda") if bias else None
    if a_mn and (m * 2) % 16:
        pytest.skip("TMA needs 16-byte row pitch")
    out = L.gemm(ah, al if passes == 3 else None, bh, bl if passes == 3 else None, m, n, k, passes, bias=bv,
                 a_mn=bool(a_mn), b_mn=bool(b_mn), split_k=split)
    ar = ah.double() + (al.double() if passes == 3 else 0)
    br = bh.double() + (bl.double() if passes == 3 else 0)
    if a_mn:
        ar = ar.t()
    if b_mn:
        br = br.t()
    ref = ar @ br.t() + (bv.double() if bias else 0)
    # the tensor core accumulates in fp32 with truncation: error grows ~ K * 2^-24 * |acc|
    tol = 4e-7 * k * passes ** 0.5 * max(1.0, ref.abs().max().item()) / 30 + 1e-5
    if passes == 3:   # the a_lo*b_lo term (2^-16 relative per product) is dropped by design
        tol += 2e-5 * k ** 0.5
    err = (out.double() - ref).abs().max().item()
    assert err < tol, (err, tol)
    if passes == 3:   # and the split planes reproduce the fp32 product
        ref32 = a.double() @ b.double().t() + (bv.double() if bias else 0)
        rel = (out.double() - ref32).abs().max().item() / ref32.abs().max().item()
        assert rel < 1e-4, rel
    # on-device cross-check kernel agrees too
    chk = L.gemm(ah, al if passes == 3 else None, bh, bl if passes == 3 else None, m, n, k, passes, bias=bv,
                 a_mn=bool(a_mn), b_mn=bool(b_mn), check=True)
    assert (chk.double() - ref).abs().max().item() < tol


def test_gemm_fused_first_sinkhorn_marginal(L):
    """the score GEMM's epilogue can accumulate u_k = sum_n exp(S_nk/eps) (first Sinkhorn pass)"""
    torch.manual_seed(5)
    for m, n, k in [(1000, 5000, 512), (333, 72, 64)]:
        a = torch.nn.functional.normalize(torch.randn(m, k, device="cuda"), dim=1)
        b = torch.nn.functional.normalize(torch.randn(n, k, device="cuda"), dim=1)
        bias = 0.02 * torch.randn(n, device="cuda")
        ah, al = planes(a)
        bh, bl = planes(b)
        u = torch.zeros(n, device="cuda")
        s = L.gemm(ah, al, bh, bl, m, n, k, 3, bias=bias, colexp=(u, 1.4426950408889634 / 0.005))
        ref = torch.exp(s.double() / 0.005).sum(0)
        torch.testing.assert_close(u.double(), ref, rtol=2e-4, atol=0)
        ws = L.SinkhornWorkspace(n, "cuda") if n % 4 == 0 else None
        if ws is not None:
            u2 = L.sinkhorn_pass(s, 200.0, True, None, None, None, m, ws)
            torch.testing.assert_close(u, u2, rtol=2e-4, atol=0)


@pytest.mark.parametrize("m,n,k,passes,f16", [
    (1000, 5000, 512, 3, False),     # odd number of 128-row tiles (8): last pair full
    (1100, 600, 512, 1, True),       # 9 m-tiles: the second CTA of the last pair has no rows
    (130, 136, 64, 1, False),        # ragged n tail inside the second B half
    (4096, 264, 200, 3, False),
    (1000, 512, 2560, 1, False),     # 256 x 512 pair tiles (nsub = 2): n % 512 == 0, long K, single pass
    (1300, 1024, 2048, 1, True),     # two 512-wide n-tiles, 11 m-tiles (ragged last pair), fp16 planes
])
def test_gemm_cta_pair(L, m, n, k, passes, f16):
    """CTA pairs (tcgen05 cta_group::2: one M=256 MMA over two SMs, half of the B tile per SM) give the
    same numbers as the 1-CTA kernel"""
    torch.manual_seed(m + n)
    a = torch.nn.functional.normalize(torch.randn(m, k, device="cuda"), dim=1)
    b = torch.nn.functional.normalize(torch.randn(n, k, device="cuda"), dim=1)
    bias = 0.02 * torch.randn(n, device="cuda")
    if f16:
        ah, al, bh, bl = a.half(), None, b.half(), None
    else:
        ah, al = planes(a)
        bh, bl = planes(b)
        if passes == 1:
            al = bl = None
    u1 = torch.zeros(n, device="cuda")
    u2 = torch.zeros(n, device="cuda")
    ref = L.gemm(ah, al, bh, bl, m, n, k, passes, bias=bias, force_m128=True, colexp=(u1, 20.0))
    got = L.gemm(ah, al, bh, bl, m, n, k, passes, bias=bias, pair=True, colexp=(u2, 20.0))
    assert torch.equal(ref, got)
    torch.testing.assert_close(u1, u2, rtol=1e-5, atol=0)
    for _ in range(3):     # repeated launches (barrier phases, cluster exit)
        got = L.gemm(ah, al, bh, bl, m, n, k, passes, bias=bias, pair=True)
    assert torch.equal(ref, got)


def test_gemm_cta_pair_mn_major_split_k(L):
    """weight-gradient shape: C[m,n] += A^T B with both operands stored [K, *] (MN-major), split-K, CTA pairs"""
    torch.manual_seed(11)
    m, n, k = 704, 512, 3000   # MN-major rows need a 16-byte pitch
    a = torch.randn(k, m, device="cuda") * 0.1
    b = torch.randn(k, n, device="cuda") * 0.1
    ah, bh = a.bfloat16(), b.bfloat16()
    ref = torch.zeros(m, n, device="cuda")
    got = torch.zeros(m, n, device="cuda")
    L.gemm(ah, None, bh, None, m, n, k, 1, out=ref, a_mn=True, b_mn=True, split_k=3, accumulate=True)
    L.gemm(ah, None, bh, None, m, n, k, 1, out=got, a_mn=True, b_mn=True, split_k=3, accumulate=True, pair=True)
    exact = ah.double().t() @ bh.double()
    tol = 4e-7 * k * exact.abs().max().item() + 1e-4
    assert (ref.double() - exact).abs().max().item() < tol
    assert (got.double() - exact).abs().max().item() < tol
    L.gemm(ah, None, bh, None, m, n, k, 1, out=got, a_mn=True, b_mn=True, split_k=5, accumulate=True, pair=True)
    assert (got.double() - 2 * exact).abs().max().item() < 2 * tol


def test_gemm_accumulate(L):
    torch.manual_seed(3)
    a = torch.randn(200, 96, device="cuda")
    b = torch.randn(64, 96, device="cuda")
    ah, al = planes(a)
    bh, bl = planes(b)
    out = torch.ones(200, 64, device="cuda")
    L.gemm(ah, al, bh, bl, 200, 64, 96, 3, out=out, accumulate=True)
    L.gemm(ah, al, bh, bl, 200, 64, 96, 3, out=out, accumulate=True)
    ref = 1 + 2 * (a.double() @ b.double().t())
    assert (out.double() - ref).abs().max().item() < 1e-3


# ------------------------------------------------------------------------------- synthesis pieces
def test_equal_linear_pixelnorm_truncate(L):
    torch.manual_seed(0)
    x = torch.randn(37, 64)
    w = torch.randn(48, 64) / 0.01
    b = torch.randn(48)
    for act in (False, True):
        ref = O.equal_linear(x, w, b, 0.01, act)
        y = L.equal_linear(x.cuda(), w.cuda(), b.cuda(), (1 / 8) * 0.01, 0.01, int(act))
        torch.testing.assert_close(y.cpu(), ref, rtol=1e-5, atol=1e-5)
    ref = x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
    torch.testing.assert_close(L.pixel_norm(x.cuda()).cpu(), ref, rtol=1e-6, atol=1e-6)
    m = torch.randn(64)
    wl = torch.randn(5, 6, 64)
    torch.testing.assert_close(L.truncate(wl.cuda(), m.cuda(), 0.7).cpu(), m + 0.7 * (wl - m), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("b,cin,cout,h,up,passes", [
    (2, 64, 64, 4, False, 3),
    (3, 128, 64, 8, False, 3),
    (1, 64, 128, 16, False, 3),
    (2, 64, 32, 20, False, 1),
    (2, 64, 64, 4, True, 3),
    (3, 128, 64, 8, True, 3),
    (1, 64, 64, 16, True, 3),
    (1, 512, 512, 32, False, 3),
    (2, 32, 16, 16, False, 3),      # BagGAN widths: channels zero-padded to the 64-wide K block
    (2, 64, 32, 8, True, 3),
    (1, 16, 16, 32, True, 3),
    (3, 32, 32, 4, False, 1),
])
def test_modconv_vs_oracle(L, b, cin, cout, h, up, passes):
    """Modulated conv (+ blur for the up path) vs the per-sample reference formulation."""
    torch.manual_seed(b * 1000 + cin + h)
    x = torch.randn(b, cin, h, h)
    style = torch.randn(b, 24)
    weight = torch.randn(1, cout, cin, 3, 3)
    mod_w = torch.randn(cin, 24)
    mod_b = 1 + 0.1 * torch.randn(cin)
    blur_k = O.make_fir_kernel([1, 3, 3, 1]) * 4
    ref = O.modulated_conv2d(x, style, weight, mod_w, mod_b, True, up, blur_k)

    s = L.equal_linear(style.cuda(), mod_w.cuda(), mod_b.cuda(), 1 / 24 ** 0.5, 1.0, 0)
    scale = 1 / (cin * 9) ** 0.5
    w_hi, w_lo, wsq = L.modconv_prepare(weight[0].contiguous().cuda(), scale)
    demod = L.modconv_demod(wsq, s)
    x_hi, x_lo = L.modulate_split(x.permute(0, 2, 3, 1).contiguous().cuda(), s, b)
    out, _, _ = L.modconv(x_hi, x_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, cout, up, passes,
                          demod=demod)
    if up:
        out, _, _ = L.blur_noise_bias_act(out, blur_k.cuda(), 1, 1, None, None, None, 0, None)
    got = out.permute(0, 3, 1, 2).cpu()
    assert got.shape == ref.shape
    tol = 2e-4 if passes == 3 else 4e-2
    torch.testing.assert_close(got, ref, rtol=tol, atol=tol)


def test_modconv_fused_epilogue(L):
    """noise + bias + lrelu*sqrt2 + next-layer modulate/split in the conv epilogue and in
    the blur kernel."""
    torch.manual_seed(5)
    b, cin, cout, h = 2, 64, 64, 8
    x = torch.randn(b, cin, h, h)
    s = 1 + 0.2 * torch.randn(b, cin)
    weight = torch.randn(cout, cin, 3, 3)
    bias = torch.randn(cout)
    strength = torch.tensor([0.37])
    nxt = 1 + 0.3 * torch.randn(b, cout)
    scale = 1 / (cin * 9) ** 0.5
    w_hi, w_lo, wsq = L.modconv_prepare(weight.cuda(), scale)
    demod = L.modconv_demod(wsq, s.cuda())
    x_hi, x_lo = L.modulate_split(x.permute(0, 2, 3, 1).contiguous().cuda(), s.cuda(), b)
    for up in (False, True):
        ho = 2 * h if up else h
        for per_sample in (False, True):
            noise = torch.randn(b if per_sample else 1, 1, ho, ho)
            wmod = scale * weight[None] * s[:, None, :, None, None]
            d = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
            wmod = wmod * d[:, :, None, None, None]
            outs = []
            for i in range(b):
                if up:
                    o = torch.nn.functional.conv_transpose2d(x[i:i + 1], wmod[i].transpose(0, 1), stride=2)
                    o = O.upfirdn2d(o, O.make_fir_kernel([1, 3, 3, 1]) * 4, pad=(1, 1))
                else:
                    o = torch.nn.functional.conv2d(x[i:i + 1], wmod[i], padding=1)
                outs.append(o)
            ref = torch.cat(outs) + strength * noise
            ref = O.fused_leaky_relu(ref, bias)
            ref_next = ref * nxt[:, :, None, None]
            if up:
                tmp, _, _ = L.modconv(x_hi, x_lo, w_hi, w_lo, cout, True, 3, demod=demod)
                out, nh, nl = L.blur_noise_bias_act(tmp, (O.make_fir_kernel([1, 3, 3, 1]) * 4).cuda(), 1, 1,
                                                    noise.cuda(), strength.cuda(), bias.cuda(), 1, nxt.cuda())
            else:
                out, nh, nl = L.modconv(x_hi, x_lo, w_hi, w_lo, cout, False, 3, demod=demod, noise=noise.cuda(),
                                        noise_strength=strength.cuda(), bias=bias.cuda(), act=1,
                                        next_style=nxt.cuda())
            torch.testing.assert_close(out.permute(0, 3, 1, 2).cpu(), ref, rtol=2e-4, atol=2e-4)
            nxt_got = (nh.float() + nl.float()).permute(0, 3, 1, 2).cpu()
            torch.testing.assert_close(nxt_got, ref_next, rtol=3e-4, atol=3e-4)


@pytest.mark.parametrize("cin,cout,h,w", [(16, 16, 32, 64), (32, 32, 20, 37), (32, 16, 16, 32), (16, 32, 5, 3),
                                          (8, 8, 33, 40)])
def test_modconv_small_vs_oracle(L, cin, cout, h, w):
    """Direct fp32 few-channel modulated conv (`gx_modconv_small`, the BagGAN 128^2 / 256^2 layers) against
    ModulatedConv2d + NoiseInjection + FusedLeakyReLU of the oracle (ref model.py:327-382), ragged tiles and image
    borders included, and against the tcgen05 path on the same layer."""
    torch.manual_seed(cin + cout + h)
    b = 3
    x = torch.randn(b, cin, h, w)
    s = 1 + 0.2 * torch.randn(b, cin)
    weight = torch.randn(cout, cin, 3, 3)
    bias = torch.randn(cout)
    strength = torch.tensor([0.37])
    nxt = 1 + 0.3 * torch.randn(b, cout)
    scale = 1 / (cin * 9) ** 0.5
    assert L.modconv_small_supported(cin, cout) and not L.modconv_small_supported(64, 64)
    w9 = L.modconv_small_weights(weight.cuda(), scale)
    _, _, wsq = L.modconv_prepare(weight.cuda(), scale)
    demod = L.modconv_demod(wsq, s.cuda())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    for per_sample in (False, True):
        noise = torch.randn(b if per_sample else 1, 1, h, w)
        wmod = scale * weight[None] * s[:, None, :, None, None]
        d = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
        wmod = wmod * d[:, :, None, None, None]
        ref = torch.cat([torch.nn.functional.conv2d(x[i:i + 1], wmod[i], padding=1) for i in range(b)])
        ref = O.fused_leaky_relu(ref + strength * noise, bias)
        out, nh, nl = L.modconv_small(x_nhwc, s.cuda(), w9, demod=demod, noise=noise.cuda(),
                                      noise_strength=strength.cuda(), bias=bias.cuda(), act=1, next_style=nxt.cuda())
        torch.testing.assert_close(out.permute(0, 3, 1, 2).cpu(), ref, rtol=2e-5, atol=2e-5)      # fp32 arithmetic
        assert nh.shape[-1] == 64 and float(nh[..., cout:].float().abs().max()) == 0.0           # zero padding
        nxt_got = (nh.float() + nl.float())[..., :cout].permute(0, 3, 1, 2).cpu()
        torch.testing.assert_close(nxt_got, ref * nxt[:, :, None, None], rtol=3e-5, atol=3e-5)
    # no demodulation / noise / bias / activation: the bare conv
    out, nh, nl = L.modconv_small(x_nhwc, s.cuda(), w9)
    assert nh is None and nl is None
    wmod = scale * weight[None] * s[:, None, :, None, None]
    ref = torch.cat([torch.nn.functional.conv2d(x[i:i + 1], wmod[i], padding=1) for i in range(b)])
    torch.testing.assert_close(out.permute(0, 3, 1, 2).cpu(), ref, rtol=2e-5, atol=2e-5)


def test_torgb(L):
    torch.manual_seed(2)
    b, c, h = 2, 64, 8
    x = torch.randn(b, c, h, h)
    w = torch.randn(3, c)
    s = torch.randn(b, c)
    bias = torch.randn(3)
    skip = torch.randn(b, 3, h, h)
    scale = 1 / c ** 0.5
    ref = torch.einsum("bchw,oc,bc->bohw", x, w * scale, s) + bias.view(1, 3, 1, 1) + skip
    got = L.torgb(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), scale, s.cuda(), bias.cuda(), skip.cuda())
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------- head pieces
def test_gather_rows_vs_oracle(L):
    torch.manual_seed(7)
    b = 2
    feats = [torch.randn(b, 8, 4, 4), torch.randn(b, 12, 8, 8), torch.randn(b, 8, 8, 8), torch.randn(b, 4, 16, 16)]
    hlen = 28   # cuts into the last map like [:, :hlen]
    hf = O.pixel_feature_vectors(feats, hlen)
    nhwc = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats]
    rows_ref, row_src, row_img = [], [], []
    for i, (ang, flip) in enumerate([(7.5, True), (-3.0, False)]):
        t = O.rotate_flip(hf[i:i + 1], ang, flip)
        perm = torch.randperm(256)
        rows_ref.append(O.sample_rows(t, perm, 100))
        from ganecdotes_b200.hfc_with_swav.engine import rotate_flip_index_map
        mp = rotate_flip_index_map(16, 16, ang, flip)
        assert torch.equal(mp, O.rotate_flip_index_map(16, 16, ang, flip))
        row_src.append(mp[perm[:100]])
        row_img.append(torch.full((100,), i))
    rows_ref = torch.cat(rows_ref)
    row_src = torch.cat(row_src).to(torch.int32).cuda()
    row_img = torch.cat(row_img).to(torch.int32).cuda()
    a_hi, a_lo, a_f = L.gather_rows(nhwc, 16, 16, hlen, row_img, row_src, 200, ld=32, want_f32=True)
    assert torch.equal(a_f[:, :hlen].cpu(), rows_ref)                 # exact: it is a gather
    assert torch.count_nonzero(a_f[:, hlen:]) == 0
    rec = (a_hi.float() + a_lo.float())[:, :hlen].cpu()
    torch.testing.assert_close(rec, rows_ref, rtol=2e-5, atol=1e-6)
    # all pixels in order (prediction path)
    a_hi, a_lo, a_f = L.gather_rows(nhwc, 16, 16, hlen, None, None, b * 256, want_f32=True)
    assert torch.equal(a_f.cpu(), hf.permute(0, 2, 3, 1).reshape(-1, hlen))


def test_gemm_fp16_planes_and_l2norm_fp16_output(L):
    torch.manual_seed(9)
    n, k, c = 700, 300, 512
    z = torch.randn(n, c).cuda()
    w = torch.nn.functional.normalize(torch.randn(k, c), dim=1).cuda()
    hi, lo, inv, zf = L.l2norm_split(z, want_lo=True, want_f16=True)
    zn = torch.nn.functional.normalize(z, dim=1)
    assert zf.dtype == torch.float16
    torch.testing.assert_close(zf.float(), zn, rtol=5e-4, atol=1e-7)     # one fp16 rounding (2^-11)
    wf = L.round_f16(w)
    assert torch.equal(wf, w.half())
    bias = torch.randn(k).cuda()
    s = L.gemm(zf, None, wf, None, n, k, c, 1, bias=bias)
    chk = L.gemm(zf, None, wf, None, n, k, c, 1, bias=bias, check=True)
    torch.testing.assert_close(s, chk, rtol=0, atol=2e-6)                  # tensor-core vs SIMT on the same planes
    ref = (zn.double() @ w.double().t() + bias.double()).float()
    err = (s - ref).abs()
    assert err.max().item() < 1e-4 and err.pow(2).mean().sqrt().item() < 2e-5, (err.max().item(),)


def test_upsample_sum_and_pool_sum(L):
    torch.manual_seed(3)
    b, c = 2, 24
    parts = [torch.randn(b, 4, 4, c), torch.randn(b, 8, 8, c), torch.randn(b, 16, 16, c)]
    ref = sum(torch.nn.functional.interpolate(p.permute(0, 3, 1, 2), size=(16, 16), mode="nearest") for p in parts)
    got = L.upsample_sum([p.cuda() for p in parts], b, 16, 16)
    torch.testing.assert_close(got.view(b, 16, 16, c).cpu(), ref.permute(0, 2, 3, 1), rtol=1e-6, atol=1e-6)
    x = torch.randn(b, 16, 16, c)
    f, hi, lo = L.pool_sum(x.cuda(), 4, 4, want_lo=True)
    pref = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 4).permute(0, 2, 3, 1) * 16
    torch.testing.assert_close(f.cpu(), pref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close((hi.float() + lo.float()).view(b, 4, 4, c).cpu(), pref, rtol=2e-5, atol=1e-5)
    # adjointness: <upsample(p), x> == <p, pool(x)>
    p = parts[0]
    up = torch.nn.functional.interpolate(p.permute(0, 3, 1, 2), size=(16, 16), mode="nearest").permute(0, 2, 3, 1)
    torch.testing.assert_close((up * x).sum(), (p * f.cpu()).sum(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("b,res,c,pads", [(2, 8, 512, (1, 1)), (3, 16, 128, (1, 1)), (2, 33, 16, (1, 1)),
                                          (1, 12, 24, (2, 1)), (2, 64, 64, (1, 1)), (1, 9, 4, (0, 3))])
def test_blur_separable_kernel_matches_2d_kernel_and_oracle(L, monkeypatch, b, res, c, pads):
    """gx_blur_sep_noise_bias_act (horizontal + vertical 4-tap pass) against the 2-D kernel and against the oracle's
    upfirdn2d + noise + bias + leaky-relu (ref model.py:166-182, 371-382, 15-43)."""
    torch.manual_seed(res + c)
    fir = (O.make_fir_kernel([1, 3, 3, 1]) * 4)
    x = torch.randn(b, res, res, c)
    ho = res + pads[0] + pads[1] - 3
    noise = torch.randn(1, ho, ho)
    strength = torch.tensor([0.37])
    bias = torch.randn(c)
    style = torch.randn(b, c)
    sep = L.separable_factors(fir.cuda())
    assert sep is not None
    outs = {}
    # the 2-D kernel maps min(C/4, 256) channel quads onto a 256-thread block: C/4 must divide 256
    modes = ("0", "1") if 256 % min(c // 4, 256) == 0 else ("1",)
    for mode in modes:
        monkeypatch.setenv("GX_BLUR_SEP", mode)
        outs[mode] = L.blur_noise_bias_act(x.cuda(), fir.cuda(), pads[0], pads[1], noise.cuda(), strength.cuda(),
                                           bias.cuda(), 1, style.cuda(), sep=sep)
    if "0" in outs:
        torch.testing.assert_close(outs["0"][0], outs["1"][0], rtol=2e-6, atol=4e-6)
        torch.testing.assert_close(outs["0"][1].float() + outs["0"][2].float(),
                                   outs["1"][1].float() + outs["1"][2].float(), rtol=3e-5, atol=1e-5)
    ref = O.upfirdn2d(x.permute(0, 3, 1, 2).contiguous(), fir, pad=pads) + strength * noise[None]
    ref = O.fused_leaky_relu(ref, bias)
    torch.testing.assert_close(outs["1"][0].cpu().permute(0, 3, 1, 2), ref, rtol=1e-5, atol=2e-6)
    planes_sum = outs["1"][1].float() + outs["1"][2].float()
    torch.testing.assert_close(planes_sum[..., :c].cpu(), (ref * style[:, :, None, None]).permute(0, 2, 3, 1),
                               rtol=3e-5, atol=1e-5)
    # no noise / bias / activation (the up-conv of ToRGB-free callers)
    monkeypatch.setenv("GX_BLUR_SEP", "1")
    o = L.blur_noise_bias_act(x.cuda(), fir.cuda(), pads[0], pads[1], None, None, None, 0, None, sep=sep)[0]
    torch.testing.assert_close(o.cpu().permute(0, 3, 1, 2),
                               O.upfirdn2d(x.permute(0, 3, 1, 2).contiguous(), fir, pad=pads), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("sizes,out_hw,c", [
    ((4, 8, 16, 32), (32, 32), 24),          # pyramid: three levels shared by 2x2 quads + one per-pixel level
    ((4, 8, 16), (32, 32), 516),             # every level shared; channel count not a multiple of 128
    ((8, 32, 32), (32, 32), 64),             # two per-pixel levels
    ((3, 6, 12), (12, 12), 8),               # odd coarse size
    ((4, 6, 12), (12, 12), 8),               # 12/4 = 3: odd factor first -> no shared level
    ((16, 8), (16, 16), 8),                  # coarse level after the fine one: both loaded per pixel
    ((16, 4, 8), (16, 16), 8),               # three levels after a per-pixel one -> per-pixel kernel
    ((5, 10), (10, 14), 8),                  # non-integer factor in x -> per-pixel kernel
])
def test_upsample_sum_quad_kernel_is_bit_identical_to_per_pixel_kernel(L, monkeypatch, sizes, out_hw, c):
    """The 2x2-quad kernel of gx_upsample_sum (levels shared by an aligned output quad are loaded once) performs
    the per-pixel kernel's additions in the same order: identical bits, and both equal the level-ordered sum of
    F.interpolate(mode='nearest') (ref swav_clustering.py:112-126)."""
    torch.manual_seed(11)
    b = 3
    oh, ow = out_hw
    parts = [torch.randn(b, s, s, c).cuda() for s in sizes]
    res = {}
    for quad in ("0", "1"):
        monkeypatch.setenv("GX_UPSUM_QUAD", quad)
        hi = torch.empty(b * oh * ow, c, dtype=torch.bfloat16, device="cuda")
        lo = torch.empty_like(hi)
        lab = torch.empty(b * oh * ow, dtype=torch.int64, device="cuda")
        res[quad] = (L.upsample_sum(parts, b, oh, ow, planes=(hi, lo), labels=lab), hi, lo, lab)
    for x, y in zip(res["0"], res["1"]):
        assert torch.equal(x, y)
    # the fused label map is the first arg-max of the sums (what a separate arg-max pass over Z gives)
    assert torch.equal(res["1"][3], L.argmax_rows(res["1"][0]))
    assert torch.equal(res["1"][3], res["1"][0].max(1)[1])
    ref = torch.zeros(b, oh, ow, c, device="cuda")
    for p in parts:
        ref += torch.nn.functional.interpolate(p.permute(0, 3, 1, 2), size=(oh, ow), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(res["1"][0].view(b, oh, ow, c), ref)
    torch.testing.assert_close(res["1"][1].float() + res["1"][2].float(), ref.view(-1, c), rtol=2e-5, atol=1e-5)


@pytest.mark.parametrize("hlen,passes", [(32, 3), (24, 3), (32, 1)])
def test_per_level_projection_equals_projection_of_upsampled_vectors(L, hlen, passes):
    """Z = Wp . concat_l(upsample(F_l)) computed per resolution (engine.project_all_pixels) against the
    oracle's upsample -> concat -> 1x1 conv (ref swav_clustering.py:108-130, :171)."""
    from ganecdotes_b200.hfc_with_swav.engine import project_all_pixels
    torch.manual_seed(5)
    b, c = 3, 64
    feats = [torch.randn(b, 8, 4, 4), torch.randn(b, 8, 8, 8), torch.randn(b, 8, 8, 8), torch.randn(b, 8, 16, 16)]
    wp = torch.randn(c, hlen) / hlen ** 0.5
    hf = O.pixel_feature_vectors(feats, hlen)                              # [b, hlen, 16, 16]
    ref = torch.einsum("bdhw,cd->bhwc", hf.double(), wp.double()).reshape(-1, c).float()
    nhwc = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats]
    wp_hi, wp_lo = L.split_planes(wp.cuda(), want_lo=passes == 3)
    z, levels = project_all_pixels(wp_hi, wp_lo, nhwc, b, 16, 16, hlen, passes)
    tol = 5e-5 if passes == 3 else 2e-2
    torch.testing.assert_close(z.cpu(), ref, rtol=tol, atol=tol)
    assert [lv["h"] for lv in levels] == [4, 8, 16][:len(levels)]
    assert sum(lv["keep"] for lv in levels) == hlen


def test_l2norm_fwd_bwd(L):
    torch.manual_seed(0)
    z = torch.randn(50, 64, requires_grad=True)
    zn = torch.nn.functional.normalize(z, p=2, dim=1)
    g = torch.randn(50, 64)
    zn.backward(g)
    hi, lo, inv = L.l2norm_split(z.detach().cuda())
    torch.testing.assert_close((hi.float() + lo.float()).cpu(), zn.detach(), rtol=2e-5, atol=1e-6)
    dh, dl = L.l2norm_bwd_split(g.cuda(), hi, lo, inv)
    torch.testing.assert_close((dh.float() + dl.float()).cpu(), z.grad, rtol=1e-4, atol=1e-5)
    w = torch.randn(30, 16)
    torch.testing.assert_close(L.normalize_rows_(w.clone().cuda()).cpu(), torch.nn.functional.normalize(w, dim=1),
                               rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n,k", [(40, 24), (333, 48), (1000, 5000), (64, 4000), (301, 8000), (77, 6144)])
def test_sinkhorn_vs_oracle(L, n, k):
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(n + k)
    s = 0.05 * torch.randn(n, k)
    ref = O.sinkhorn_knopp(s.double(), 10, 0.005).float()
    ws = L.SinkhornWorkspace(k, "cuda")
    la = E.sinkhorn_log_a(s.cuda(), 10, 0.005, ws, n)
    q = L.sinkhorn_q(s.cuda(), 1 / 0.005, la)
    torch.testing.assert_close(q.cpu(), ref, rtol=2e-3, atol=1e-9)
    assert abs(q.sum(1).max().item() - 1) < 1e-4


@pytest.mark.parametrize("n,k,std", [(40, 24, 0.05), (333, 48, 0.05), (1000, 5000, 0.05), (257, 4000, 0.02),
                                     (301, 8000, 0.05), (77, 6144, 0.05), (513, 1500, 0.05), (64, 520, 0.05)])
def test_sinkhorn_cached_passes_vs_oracle(L, n, k, std):
    """gx_sinkhorn_pass_cached: iteration 2 (engine.CACHE16_WRITE_IT) stores its row-normalised terms as a 16-bit plane
    (row pitch rounded up to 8 columns: k = 1500 -> 1504), iterations 3.. stream that plane instead of S.  Only log a comes out of the passes;
    the codes softmax_k(S/eps + log a) are evaluated from the fp32 scores and stay within the tolerance of the fp32
    passes (rtol 2e-3 against the fp64 oracle; the cache itself moves them by < 1e-3)."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(n + k)
    s = (std * torch.randn(n, k)).cuda()
    ref = O.sinkhorn_knopp(s.cpu().double(), 10, 0.005).float()
    ws = L.SinkhornWorkspace(k, "cuda")
    la32 = E.sinkhorn_log_a(s, 10, 0.005, ws, n)
    la16 = E.sinkhorn_log_a(s, 10, 0.005, ws, n, cache16=True)
    assert not torch.equal(la16, la32)                       # the cached kernels did run
    torch.testing.assert_close(la16, la32, rtol=0, atol=1e-3)
    # ... and computed what the scheme says (fp64 restatement with the same fp16 rounding of the plane): the distance
    # to the fp32 passes above is the cache's rounding, not the kernels'
    # (a term that sits on an fp16 rounding boundary can fall the other way in fp32 arithmetic - one ulp = 1e-3 of
    # that term - which shows in the columns that a single row dominates: the median distance is fp32 rounding, 1-2 %
    # of the columns may differ by more than 1e-4; against the fp32 passes the median distance is ~1e-4)
    d = (la16.cpu() - O.sinkhorn_log_a_cached16(s.cpu(), 10, 0.005).float()).abs()
    assert d.median().item() < 2e-5 and d.max().item() < 2e-3 and (d > 1e-4).sum().item() <= max(2, 0.02 * k), \
        (d.median().item(), d.max().item(), (d > 1e-4).sum().item())
    q = L.sinkhorn_q(s, 1 / 0.005, la16)
    torch.testing.assert_close(q.cpu(), ref, rtol=2e-3, atol=1e-9)
    assert abs(q.sum(1).max().item() - 1) < 1e-4
    # a second call re-uses the workspace's plane (and a smaller problem a slice of it)
    la16b = E.sinkhorn_log_a(s, 10, 0.005, ws, n, cache16=True)
    assert torch.equal(la16b, la16)
    m = max(n // 2, 8)
    la_h = E.sinkhorn_log_a(s[:m], 10, 0.005, ws, m, cache16=True)
    torch.testing.assert_close(la_h, E.sinkhorn_log_a(s[:m], 10, 0.005, ws, m), rtol=0, atol=1e-3)


def test_sinkhorn_cached_passes_image_pdf_and_fused_first_marginals(L):
    """Non-uniform marginals (source_pdf == 'image'): the engine keeps the fp32 passes even when the cache is asked
    for (an empty histogram bin gives a prototype 1e-9 counts of target mass: its column of the row-normalised plane
    is below the fp16 range) - here 500 bins over 400 pixels, most of them empty.  The kernel itself handles mildly
    non-uniform marginals (driven directly below).  And the cached passes with the first marginals taken from the score
    GEMM's epilogue (u_first), as the training step calls them."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(1)
    n, k = 1500, 500
    sc = 0.05 * torch.randn(n, k)
    r, c = O.image_marginals(torch.rand(1, 20, 20), k, n)
    ref = O.sinkhorn_knopp(sc.double(), 10, 0.005, r.double(), c.double()).float()
    ws = L.SinkhornWorkspace(k, "cuda")
    la = E.sinkhorn_log_a(sc.cuda(), 10, 0.005, ws, n, None, r.cuda(), c.cuda(), cache16=True)
    assert torch.equal(la, E.sinkhorn_log_a(sc.cuda(), 10, 0.005, ws, n, None, r.cuda(), c.cuda()))
    q = L.sinkhorn_q(sc.cuda(), 1 / 0.005, la)
    torch.testing.assert_close(q.cpu(), ref, rtol=2e-3, atol=1e-9)
    # the kernel with marginals that vary by a factor of a few (no empty bins): u after iteration 2 through the cache
    n, k = 96, 32
    sc = (0.08 * torch.randn(n, k)).cuda()
    r, c = O.image_marginals(torch.rand(1, 12, 12), k, n)
    r, c = r.cuda(), c.cuda()
    ws = L.SinkhornWorkspace(k, "cuda")
    u0 = L.sinkhorn_pass(sc, 200.0, True, None, r, c, n, ws).clone()
    u1 = L.sinkhorn_pass(sc, 200.0, False, u0, r, c, n, ws).clone()
    u2 = L.sinkhorn_pass(sc, 200.0, False, u1, r, c, n, ws).clone()
    cache = ws.cache16(0, n)
    u1c = torch.empty_like(u1)
    L.sinkhorn_reduce(ws.partials, L.sinkhorn_pass_cached_parts(sc, 200.0, u0, r, c, n, ws, cache, True), k, u1c)
    torch.testing.assert_close(u1c, u1, rtol=1e-5, atol=0)
    u2c = torch.empty_like(u1)
    L.sinkhorn_reduce(ws.partials, L.sinkhorn_pass_cached_parts(sc, 200.0, u1c, r, c, n, ws, cache, False), k, u2c)
    torch.testing.assert_close(u2c, u2, rtol=1e-3, atol=0)
    n, k = 700, 5000
    s = (0.05 * torch.randn(n, k)).cuda()
    u0 = torch.exp(s.double() / 0.005).sum(0).float()
    la_u = E.sinkhorn_log_a(s, 10, 0.005, ws if ws.k == k else L.SinkhornWorkspace(k, "cuda"), n, u_first=u0,
                            cache16=True)
    q = L.sinkhorn_q(s, 1 / 0.005, la_u)
    torch.testing.assert_close(q.cpu(), O.sinkhorn_knopp(s.cpu().double(), 10, 0.005).float(), rtol=2e-3, atol=1e-9)


def test_sinkhorn_golden_and_image_pdf(L):
    from ganecdotes_b200.hfc_with_swav import engine as E
    g = load("swav")
    s = g["sk_scores_s"].cuda()
    ws = L.SinkhornWorkspace(s.shape[1], "cuda")
    la = E.sinkhorn_log_a(s, 10, 0.005, ws, s.shape[0])
    q = L.sinkhorn_q(s, 1 / 0.005, la)
    torch.testing.assert_close(q.cpu(), g["sk_q_s"], rtol=2e-3, atol=1e-9)
    # non-uniform marginals (source_pdf == 'image')
    torch.manual_seed(1)
    n, k = 96, 32
    sc = 0.08 * torch.randn(n, k)
    img = torch.rand(1, 12, 12)
    r, c = O.image_marginals(img, k, n)
    ref = O.sinkhorn_knopp(sc.double(), 10, 0.005, r.double(), c.double()).float()
    ws = L.SinkhornWorkspace(k, "cuda")
    la = E.sinkhorn_log_a(sc.cuda(), 10, 0.005, ws, n, None, r.cuda(), c.cuda())
    q = L.sinkhorn_q(sc.cuda(), 1 / 0.005, la)
    torch.testing.assert_close(q.cpu(), ref, rtol=2e-3, atol=1e-9)


@pytest.mark.parametrize("n,k,temp", [(40, 24, 0.01), (257, 5000, 0.01), (300, 1500, 0.01), (1000, 3000, 0.01),
                                      (90, 8192, 0.01), (257, 5000, 0.005), (257, 5000, 0.013), (64, 520, 0.02)])
def test_swav_loss_fwd_bwd_vs_oracle(L, n, k, temp):
    """temp / eps = 2 and 1 take the kernel that derives the Sinkhorn code numerators from the softmax(S/T)
    numerators (exp(S/eps + log a) = exp(S/T)^rho * a); other ratios the general kernel."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(n)
    s_s = (0.05 * torch.randn(n, k)).requires_grad_(True)
    s_t = (0.05 * torch.randn(n, k)).requires_grad_(True)
    q_s = O.sinkhorn_knopp(s_s.detach(), 10, 0.005)
    q_t = O.sinkhorn_knopp(s_t.detach(), 10, 0.005)
    loss = O.swapped_prediction_loss(s_s / temp, s_t / temp, q_s, q_t)
    loss.backward()
    ws = L.SinkhornWorkspace(k, "cuda")
    la_s = E.sinkhorn_log_a(s_s.detach().cuda(), 10, 0.005, ws, n)
    la_t = E.sinkhorn_log_a(s_t.detach().cuda(), 10, 0.005, ws, n)
    parts, ds_s, ds_t, db, f32 = L.swav_loss(s_s.detach().cuda(), s_t.detach().cuda(), 200.0, 1.0 / temp, la_s, la_t,
                                              1.0 / n, want_lo=True, want_f32=True)
    got = parts.sum().item() / n
    assert abs(got - loss.item()) < 2e-4 * abs(loss.item()) + 1e-5, (got, loss.item())
    gmax = s_s.grad.abs().max().item()
    assert (f32[0].cpu() - s_s.grad).abs().max().item() < 3e-3 * gmax
    assert (f32[1].cpu() - s_t.grad).abs().max().item() < 3e-3 * gmax
    rec = ds_s[0].float() + ds_s[1].float()
    assert (rec - f32[0]).abs().max().item() < 1e-4 * gmax
    assert (ds_t[0].float() - f32[1]).abs().max().item() < 1e-2 * gmax          # bf16 plane alone
    torch.testing.assert_close(db.cpu(), (s_s.grad + s_t.grad).sum(0), rtol=1e-2, atol=3e-3 * gmax)


def test_larc_sgd_vs_oracle(L):
    torch.manual_seed(0)
    p = torch.randn(1000, 33)
    bufs = [None]
    ps = [p]
    pd = p.clone().cuda()
    buf = torch.zeros_like(pd)
    norms = L.larc_scratch("cuda")
    for step in range(3):
        g = torch.randn(1000, 33) * 0.01
        ps, bufs = O.larc_sgd_step(ps, [g], bufs, 0.01, 0.9, 0.01)
        L.larc_sgd_(pd, g.cuda(), buf, 0.01, 0.9, 0.01, 0.0, 1e-8, step == 0, norms)
        torch.testing.assert_close(pd.cpu(), ps[0], rtol=1e-5, atol=1e-6)


def test_argmax_and_kmeans(L):
    torch.manual_seed(0)
    x = torch.randn(1000, 512)
    x[5, 7] = x[5, 300] = 9.0            # tie -> first index
    lab = L.argmax_rows(x.cuda())
    assert lab.dtype == torch.int64
    assert torch.equal(lab.cpu(), x.max(1)[1])
    assert lab[5].item() == 7
    xv = torch.randn(700, 96)
    c = torch.randn(16, 96)
    got = L.kmeans_assign(xv.cuda(), c.cuda())
    assert got.dtype == torch.int32
    assert torch.equal(got.cpu(), O.kmeans_assign(xv, c))
    got2 = L.kmeans_assign(xv[:, :64].contiguous().cuda(), c.cuda(), xv[:, 64:].contiguous().cuda())
    assert torch.equal(got2, got)
    # exact ties -> first centre
    c2 = torch.stack([c[0], c[1], c[1]])
    assert set(L.kmeans_assign(c2.cuda(), c2.cuda()).cpu().tolist()) <= {0, 1}


@pytest.mark.parametrize("n,c1,c2,k", [(5000, 64, 32, 5), (4099, 48, 0, 20), (6001, 256, 256, 50), (4096, 16, 16, 64),
                                       (4100, 512, 512, 3), (40000, 512, 512, 32), (777, 64, 0, 17), (100, 128, 64, 64)])
def test_kmeans_fused_kernel_vs_oracle(L, n, c1, c2, k):
    """gx_kmeans_assign_mma (k <= 64, channels % 16 == 0): ragged row counts, k that is not a multiple of 8, one or two
    feature maps, more row tiles than resident warps; labels against the oracle's fp32 assignment (they may differ only at near-ties) and against the
    GEMM route; exact ties go to the first centre"""
    torch.manual_seed(n + k)
    x = torch.randn(n, c1 + c2) + 0.3
    cen = torch.randn(k, c1 + c2) + 0.3
    x[7] = cen[k - 1]                     # a row that IS a centre
    xg = x.cuda()
    a = xg[:, :c1].contiguous()
    b2 = xg[:, c1:].contiguous() if c2 else None
    lab = L.kmeans_assign(a, cen.cuda(), b2, tensor=True)
    assert lab.dtype == torch.int32 and lab.shape == (n,)
    assert lab[7].item() == k - 1
    d = torch.cdist(x.double(), cen.double()) ** 2
    ref = O.kmeans_assign(x, cen)
    for got in (lab, L.kmeans_assign(a, cen.cuda(), b2, tensor="gemm")):
        mism = got.cpu().long() != ref.long()
        if mism.any():
            top2 = d.topk(2, dim=1, largest=False).values
            assert ((top2[:, 1] - top2[:, 0])[mism] < 1e-4 * top2[:, 0][mism]).all()
        assert mism.float().mean().item() < 2e-3
    # duplicated centres: the first of the two wins
    cen2 = torch.cat([cen[:1], cen[:1], cen[1:]])[:k].contiguous().cuda()
    lab2 = L.kmeans_assign(a, cen2, b2, tensor=True)
    assert not (lab2 == 1).any() or k == 1


def test_kmeans_layer_maps_vs_oracle(L):
    """config 5: per-layer assignment + one-hot NEAREST maps on generator features"""
    from ganecdotes_b200.hfc_kmeans.hfc_kmeans_clustering import FlatKMeansAssign
    torch.manual_seed(4)
    b = 2
    shapes = [(32, 4)] + [(32, 8), (32, 8), (16, 16), (16, 16)]
    feats = [torch.randn(b, c, r, r) for c, r in shapes]
    ks = [4, 8]
    centers = [torch.randn(ks[0], 64), torch.randn(ks[1], 32)]
    pairs = [torch.cat([feats[2 * n + 1], feats[2 * n + 2]], 1) for n in range(2)]
    ref_maps, ref_labels = O.kmeans_layer_maps(pairs, centers, 32)
    ref_maps = (ref_maps + 1) / 2          # the oracle helper returns the {-1,+1} encoding
    km = FlatKMeansAssign(centers, out_size=32)
    maps, labels = km.predict([f.cuda().contiguous(memory_format=torch.channels_last) for f in feats])
    for a, r in zip(labels, ref_labels):
        assert torch.equal(a.cpu(), r)
    assert torch.equal(maps.cpu(), ref_maps)


def test_split_planes(L):
    torch.manual_seed(0)
    x = torch.randn(70, 45, device="cuda")
    hi, lo = L.split_planes(x)
    assert (hi.float() + lo.float() - x).abs().max().item() < 2e-5 * x.abs().max().item()
    hi_t, lo_t = L.split_planes(x, transpose=True)
    assert torch.equal(hi_t, hi.t().contiguous()) and torch.equal(lo_t, lo.t().contiguous())


def test_kmeans_fit_statistical_parity_with_sklearn(L):
    """Lloyd + greedy k-means++ on the GPU.  scikit-learn's fit depends on its version and
    RNG, so parity is statistical: the inertia of the fit is within a few percent of scikit-learn's on the same
    data, and the structural properties of a Lloyd fixed point hold exactly."""
    from sklearn.cluster import KMeans
    from ganecdotes_b200.hfc_kmeans.hfc_kmeans_clustering import kmeans_fit
    rs = np.random.RandomState(0)
    k, c, n = 16, 96, 6000
    true_c = rs.randn(k, c) * 3
    x = (true_c[rs.randint(k, size=n)] + rs.randn(n, c)).astype(np.float32)
    ref = KMeans(n_clusters=k, n_init=1, random_state=0).fit(x)
    xg = torch.from_numpy(x).cuda()
    cen, labels, inertia, n_iter = kmeans_fit(xg[:, :64].contiguous(), k, seed=0, x2=xg[:, 64:].contiguous())
    assert cen.shape == (k, c) and labels.dtype == torch.int32 and n_iter >= 1
    assert abs(inertia - ref.inertia_) < 0.03 * ref.inertia_, (inertia, ref.inertia_)
    # fixed-point properties: labels are the nearest centres, centres are the means of their points
    d = torch.cdist(xg.double(), cen.double())
    assert (labels.long() == d.argmin(1)).float().mean().item() > 0.999
    for j in range(k):
        m = labels == j
        if m.any():
            torch.testing.assert_close(cen[j], xg[m].mean(0), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(torch.tensor(inertia), (d.min(1).values ** 2).sum().float().cpu(), rtol=1e-3, atol=0)


def test_upfirdn2d_and_fused_lrelu_backward_and_double_backward(L):
    """the ops are differentiable to second order like the reference's autograd pairs
    (lib/gan/optim/upfirdn2d.py:17-143, fused_act.py:27-167); checked against torch autograd of the native
    formulas (the oracle's upfirdn2d / fused_leaky_relu)."""
    from ganecdotes_b200.stylegan2.op import upfirdn2d, fused_leaky_relu
    torch.manual_seed(0)
    k = torch.tensor([1., 3., 3., 1.])
    k2 = (k[:, None] * k[None, :] / 64).cuda()
    for up, down, pad in [(1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1)), (2, 2, (1, 2, 0, 1))]:
        x = torch.randn(2, 3, 9, 11, device="cuda", requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        y = upfirdn2d(x, k2, up=up, down=down, pad=pad)
        yr = O.upfirdn2d(xr, k2, up=(up, up), down=(down, down), pad=pad if len(pad) == 4 else (pad[0], pad[1], pad[0], pad[1]))
        torch.testing.assert_close(y, yr, rtol=1e-5, atol=1e-5)
        g = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, g, create_graph=True)
        (gxr,) = torch.autograd.grad(yr, xr, g, create_graph=True)
        torch.testing.assert_close(gx, gxr, rtol=1e-4, atol=1e-5)
        # double backward: d/dg <gx, v> = upfirdn(v)
        v = torch.randn_like(gx)
        g2 = g.clone().requires_grad_(True)
        (gx2,) = torch.autograd.grad(upfirdn2d(x, k2, up=up, down=down, pad=pad), x, g2, create_graph=True)
        (gg,) = torch.autograd.grad(gx2, g2, v)
        torch.testing.assert_close(gg, upfirdn2d(v, k2, up=up, down=down, pad=pad).detach(), rtol=1e-4, atol=1e-5)
    # fused leaky relu
    x = torch.randn(4, 6, 5, 7, device="cuda", requires_grad=True)
    b = torch.randn(6, device="cuda", requires_grad=True)
    xr, br = x.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = fused_leaky_relu(x, b)
    yr = O.fused_leaky_relu(xr, br)
    torch.testing.assert_close(y, yr, rtol=1e-6, atol=1e-6)
    g = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, (x, b), g, create_graph=True)
    gxr, gbr = torch.autograd.grad(yr, (xr, br), g)
    torch.testing.assert_close(gx, gxr, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gb, gbr, rtol=1e-4, atol=1e-5)
    # second order: gx is linear in g with the same slopes
    g2 = g.clone().requires_grad_(True)
    gx2, gb2 = torch.autograd.grad(fused_leaky_relu(x, b), (x, b), g2, create_graph=True)
    v = torch.randn_like(gx2)
    (gg,) = torch.autograd.grad(gx2, g2, v)
    slope = torch.where(y.detach() > 0, 1.0, 0.2) * 2 ** 0.5
    torch.testing.assert_close(gg, v * slope, rtol=1e-5, atol=1e-6)


def test_bilinear_upsample_sum_and_its_adjoint(L):
    """hf_interp='bilinear' (ref swav_clustering.py:112-126): upsample_sum with F.interpolate(mode='bilinear',
    align_corners=False) semantics, and `pool_bilinear_adjoint` as its exact adjoint (<U p, x> = <p, U^T x>)."""
    torch.manual_seed(4)
    b, c = 2, 24
    parts = [torch.randn(b, 4, 4, c), torch.randn(b, 8, 8, c), torch.randn(b, 32, 32, c)]
    ref = sum(torch.nn.functional.interpolate(p.permute(0, 3, 1, 2), size=(32, 32), mode="bilinear",
                                              align_corners=False) for p in parts)
    got = L.upsample_sum([p.cuda() for p in parts], b, 32, 32, bilinear=True)
    torch.testing.assert_close(got.view(b, 32, 32, c).cpu(), ref.permute(0, 2, 3, 1), rtol=1e-5, atol=1e-5)
    x = torch.randn(b, 32, 32, c)
    for p in parts[:2]:
        h = p.shape[1]
        up = torch.nn.functional.interpolate(p.permute(0, 3, 1, 2), size=(32, 32), mode="bilinear",
                                             align_corners=False).permute(0, 2, 3, 1)
        adj = L.pool_bilinear_adjoint(x.cuda(), h, h).cpu()
        torch.testing.assert_close((up.double() * x.double()).sum(), (p.double() * adj.double()).sum(),
                                   rtol=1e-5, atol=1e-4)
        # and against autograd of F.interpolate
        pr = p.clone().requires_grad_(True)
        torch.nn.functional.interpolate(pr.permute(0, 3, 1, 2), size=(32, 32), mode="bilinear",
                                        align_corners=False).backward(x.permute(0, 3, 1, 2))
        torch.testing.assert_close(adj, pr.grad, rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------------------
# round 2: exchange over peer memory, index bookkeeping, fused W+ construction
# ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("cached", [False, True])
@pytest.mark.parametrize("world,n,k", [(2, 600, 48), (3, 999, 5000), (4, 64, 4000), (2, 300, 8000)])
def test_sinkhorn_ll_exchange_simulated_ranks(L, world, n, k, cached):
    """Distributed Sinkhorn through the tagged-word exchange (gx_sinkhorn_reduce_send + the receiving prologue
    of gx_sinkhorn_pass / gx_sinkhorn_log_a) with `world` endpoints simulated in one process: the rows of S are
    split into `world` shards, every shard's pass pushes its marginals into all buffers, the next pass of every
    shard receives the sum.  Result == the oracle's Sinkhorn on the whole matrix, and == the single-rank kernel
    path bit for bit in log a up to the summation order of the shards."""
    from ganecdotes_b200.hfc_with_swav import engine as E
    torch.manual_seed(world * 1000 + n + k)
    s = (0.05 * torch.randn(n, k)).cuda()
    bounds = [n * r // world for r in range(world + 1)]
    shards = [s[bounds[r]:bounds[r + 1]].contiguous() for r in range(world)]
    ends = L.LLExchange.simulated(world, k, "cuda")
    ws = L.SinkhornWorkspace(k, "cuda")
    inv_eps, niters = 1 / 0.005, 10
    try:
        for it in range(niters):
            for ch in (0, 1):                                   # two independent chains on two channels
                descs = []
                for r, e in enumerate(ends):                    # every endpoint sends ...
                    u_ll = e.last(ch) if it > 0 else None       # ... after receiving the previous exchange
                    if cached and it >= E.CACHE16_WRITE_IT:      # every (chain, shard) has its own 16-bit plane
                        nparts = L.sinkhorn_pass_cached_parts(shards[r], inv_eps, None, None, None, n, ws,
                                                              ws.cache16((ch, r), shards[r].shape[0]),
                                                              it == E.CACHE16_WRITE_IT, u_ll=u_ll,
                                                              reverse=(it & 1) == 1)
                    else:
                        nparts = L.sinkhorn_pass_parts(shards[r], inv_eps, it == 0, None, None, None, n, ws, u_ll=u_ll,
                                                       reverse=(it & 1) == 1)
                    descs.append(e.next_send(ch))
                    # NB: a receive only completes once ALL sends of its exchange are on the stream: sends of
                    # exchange `it` are issued below, receives of exchange `it` at iteration it + 1
                    L.sinkhorn_reduce_send(ws.partials, nparts, k, descs[-1])
        las = [[L.sinkhorn_log_a(None, None, u_ll=e.last(ch), k=k, device="cuda") for e in ends] for ch in (0, 1)]
        torch.cuda.synchronize()
        for e in ends:
            e.check()
        for ch in (0, 1):
            for r in range(1, world):                           # every rank computes the bit-identical sum
                assert torch.equal(las[ch][0], las[ch][r])
        assert torch.equal(las[0][0], las[1][0])
        la_one = E.sinkhorn_log_a(s, niters, 0.005, ws, n)
        torch.testing.assert_close(las[0][0], la_one, rtol=0, atol=1e-3 if cached else 2e-4)
        q = L.sinkhorn_q(s, inv_eps, las[0][0])
        ref = O.sinkhorn_knopp(s.cpu().double(), niters, 0.005).float()
        torch.testing.assert_close(q.cpu(), ref, rtol=2e-3, atol=1e-9)
        u = L.ll_recv_sum(ends[0].last(0), k, "cuda")           # stand-alone receive kernel == the fused prologue
        assert torch.isfinite(u).all() and (u > 0).all()
    finally:
        ends[0].close()


def test_sinkhorn_reverse_pass_is_the_same_sum(L):
    """serpentine passes: streaming the rows last-to-first changes only the summation order"""
    torch.manual_seed(5)
    n, k = 1237, 5000
    s = (0.05 * torch.randn(n, k)).cuda()
    ws = L.SinkhornWorkspace(k, "cuda")
    u0 = L.sinkhorn_pass(s, 200.0, True, None, None, None, n, ws).clone()
    uf = L.sinkhorn_pass(s, 200.0, False, u0, None, None, n, ws, reverse=False).clone()
    ur = L.sinkhorn_pass(s, 200.0, False, u0, None, None, n, ws, reverse=True).clone()
    torch.testing.assert_close(uf, ur, rtol=2e-5, atol=0)
    ur0 = L.sinkhorn_pass(s, 200.0, True, None, None, None, n, ws, reverse=True)
    torch.testing.assert_close(ur0, u0, rtol=2e-5, atol=0)


@pytest.mark.parametrize("b,hw,patches,n", [(1, 256, 2, 100), (3, 4096, 5, 3000), (8, 65536, 5, 20000)])
def test_pixel_segments_vs_torch(L, b, hw, patches, n):
    """gx_pixel_segments (counting sort) against a stable torch argsort: ridx, CSR offsets, and the sample order
    inside every segment (ascending sample index)."""
    g = torch.Generator().manual_seed(b + hw)
    row_src = torch.stack([torch.cat([torch.randperm(hw, generator=g)[:n] for _ in range(b)]) for _ in range(patches)])
    fill = torch.rand(row_src.shape, generator=g) < 0.07
    row_src[fill] = -1                                                           # rotation fill
    row_img = torch.arange(b).repeat_interleave(n)
    ridx, order, seg_off = L.pixel_segments(row_src.int().cuda(), row_img.int().cuda(), hw, b * hw)
    ref = torch.where(row_src >= 0, row_img.unsqueeze(0) * hw + row_src, torch.full_like(row_src, -1))
    assert torch.equal(ridx.cpu().long(), ref)
    keys = ref.flatten()
    valid = (keys >= 0).nonzero().flatten()
    ref_order = valid[torch.argsort(keys[valid], stable=True)]
    counts = torch.bincount(keys[valid], minlength=b * hw)
    ref_off = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    assert torch.equal(seg_off.cpu().long(), ref_off)
    assert torch.equal(order.cpu().long()[:valid.numel()], ref_order)


@pytest.mark.parametrize("c", [512, 64, 520, 20])
@pytest.mark.parametrize("bf16_rows", [False, True])
def test_segment_sum_rows_is_the_sequential_sum(L, c, bf16_rows):
    """gx_segment_sum_rows: out[seg] = rows[order[r0]] + rows[order[r0 + 1]] + ... added in that order in fp32 (empty
    segments give zeros; segments longer than a warp; c = 20 takes the 8-byte-load path of the bf16 rows).  Bit-exact
    against a sequential fp32 sum; the bf16 planes are the split of that sum."""
    g = torch.Generator().manual_seed(c + int(bf16_rows))
    nrows, nseg = 3000, 400
    rows = torch.randn(nrows, c, generator=g)
    if bf16_rows:
        rows = rows.bfloat16()
    lens = torch.randint(0, 6, (nseg,), generator=g)
    lens[7], lens[100] = 70, 33                                     # more rows than lanes
    seg_off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)]).int()
    order = torch.randint(0, nrows, (int(seg_off[-1]),), generator=g).int()
    ref = torch.zeros(nseg, c)
    rf = rows.float()
    for sgi in range(nseg):
        acc = torch.zeros(c)
        for r in range(int(seg_off[sgi]), int(seg_off[sgi + 1])):
            acc = acc + rf[order[r]]
        ref[sgi] = acc
    hi, lo, f = L.segment_sum_rows(rows.cuda(), order.cuda(), seg_off.cuda(), nseg, want_lo=True, want_f32=True)
    assert torch.equal(f.cpu(), ref)
    torch.testing.assert_close((hi.float() + lo.float()).cpu(), ref, rtol=2e-5, atol=1e-6)
    hi1, _, _ = L.segment_sum_rows(rows.cuda(), order.cuda(), seg_off.cuda(), nseg)
    assert torch.equal(hi1, hi)


def test_view_wplus_kernel_matches_host_construction(L):
    """gx_view_wplus (both views, one launch) == engine.view_wplus (reference op order, ref swav_clustering.py:593-640)"""
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator
    sd = O.init_generator_state(16, 64, 2, 7)
    gen = Generator(16, 64, 2)
    gen.load_state_dict(sd, strict=True)
    gen = gen.cuda()
    torch.manual_seed(4)
    b, nl, pstd = 3, 3, [1.0, 0.5, 0.25]
    mean = torch.randn(1, 64).cuda()
    z = torch.randn(b, 64)
    views = [E.ViewDraws(layer_no=[int(v) for v in torch.randint(0, nl, (b,))], pert_z=torch.randn(b, 2 * nl, 64),
                         angle=[0.0] * b, flip=[False] * b) for _ in range(2)]
    w = gen.style(z.cuda())
    for psi in (0.7, 1.0):
        ref = torch.cat([E.view_wplus(gen, w, mean, psi, v, pstd) for v in views])
        rows, lno, sig = [], [], []
        for v in views:
            for i, l in enumerate(v.layer_no):
                rows += [v.pert_z[i, 2 * l], v.pert_z[i, 2 * l + 1]]
                lno.append(l)
                sig.append(pstd[l])
        noise_w = gen.style(torch.stack(rows).cuda())
        got = L.view_wplus(w, noise_w, torch.tensor(lno, dtype=torch.int32).cuda(), torch.tensor(sig).cuda(),
                           mean.reshape(-1), psi, gen.n_latent)
        torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6)


def test_colsum_scale_accumulate(L):
    torch.manual_seed(0)
    parts = torch.randn(37, 101).cuda()
    out = torch.full((101,), 2.0).cuda()
    L.colsum(parts, 37, 101, out, scale=0.5, accumulate=True)
    torch.testing.assert_close(out.cpu(), 2.0 + 0.5 * parts.cpu().double().sum(0).float(), rtol=1e-5, atol=1e-6)
    L.colsum(parts, 37, 101, out, scale=1.0, accumulate=False)
    torch.testing.assert_close(out.cpu(), parts.cpu().double().sum(0).float(), rtol=1e-5, atol=1e-6)
    one = torch.zeros(1).cuda()
    L.colsum(parts.view(-1, 1), parts.numel(), 1, one, scale=2.0)
    assert abs(one.item() - 2.0 * parts.double().sum().item()) < 1e-3


def test_native_ops_match_the_reference_cuda_kernels(L):
    """gx_upfirdn2d / gx_fused_bias_act against the reference's OWN CUDA kernels (lib/gan/optim/upfirdn2d_kernel.cu,
    fused_bias_act_kernel.cu), compiled unmodified for sm_100a into oracle/_ref/ by oracle/build_ref.py: every
    (up, down, pad) family the reference's kernel table covers, and all six act/grad modes of fused_bias_act."""
    from oracle import build_ref
    up, fb = build_ref.load("ref_upfirdn2d"), build_ref.load("ref_fused_bias_act")
    if up is None or fb is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py in the build container)")
    g = torch.Generator().manual_seed(0)
    k4 = torch.tensor([1., 3., 3., 1.])
    k4 = (k4[None, :] * k4[:, None]) / 64
    k3 = torch.randn(3, 3, generator=g)
    k6 = torch.randn(6, 6, generator=g)
    cases = [((3, 17, 19, 5), k4, (1, 1), (1, 1), (2, 1, 2, 1)),       # blur
             ((2, 16, 16, 8), k4 * 4, (2, 2), (1, 1), (2, 1, 2, 1)),   # upsample
             ((2, 16, 16, 3), k4, (1, 1), (2, 2), (1, 1, 1, 1)),       # downsample
             ((1, 9, 11, 2), k3, (1, 1), (1, 1), (1, 1, 1, 1)),
             ((2, 13, 7, 4), k6, (2, 2), (1, 1), (3, 2, 3, 2)),
             ((1, 8, 8, 1), k3, (2, 1), (1, 2), (2, 0, -1, 3)),        # generic path, a negative pad crops
             # planar (minor == 1) tiled kernel: tiles with ragged edges, odd pads, 3x3 and 2x2 filters
             ((5, 67, 70, 1), k4, (1, 1), (1, 1), (2, 1, 2, 1)),
             ((3, 257, 257, 1), k4, (1, 1), (1, 1), (1, 1, 1, 1)),
             ((3, 40, 45, 1), k4 * 4, (2, 2), (1, 1), (2, 1, 2, 1)),
             ((2, 33, 64, 1), k3, (2, 2), (1, 1), (1, 2, 0, 3)),
             ((4, 70, 66, 1), k4, (1, 1), (2, 2), (1, 1, 1, 1)),
             ((2, 90, 131, 1), k3, (1, 1), (2, 2), (2, 0, 1, 1)),
             ((2, 64, 64, 1), torch.randn(2, 2, generator=g), (1, 1), (1, 1), (0, 1, 1, 0)),
             ((2, 48, 80, 1), k4, (1, 1), (1, 1), (3, -2, -1, 4))]
    for shape, k, (ux, uy), (dx, dy), (px0, px1, py0, py1) in cases:
        x = torch.randn(*shape, generator=g).cuda()
        kc = k.contiguous().cuda()
        ref = up.upfirdn2d(x, kc, ux, uy, dx, dy, px0, px1, py0, py1)
        got = L.upfirdn2d_raw(x, kc, ux, uy, dx, dy, px0, px1, py0, py1)
        assert got.shape == ref.shape
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    x = torch.randn(4, 6, 9, 9, generator=g).cuda()
    b = torch.randn(6, generator=g).cuda()
    refer = torch.randn(4, 6, 9, 9, generator=g).cuda()
    empty = torch.empty(0, device="cuda")
    for act in (1, 3):
        for grad in (0, 1, 2):
            r_in = refer if grad > 0 else empty
            ref = fb.fused_bias_act(x, b if grad == 0 else empty, r_in, act, grad, 0.2, 1.4142135)
            got = L.fused_bias_act_raw(x, b if grad == 0 else None, refer if grad > 0 else None, act, grad, 0.2, 1.4142135)
            torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("up,down,pad,shape,taps", [
    (1, 1, (2, 1), (2, 3, 67, 70), [1, 3, 3, 1]), (2, 1, (2, 1), (1, 4, 40, 45), [1, 3, 3, 1]),
    (1, 2, (1, 1), (2, 2, 70, 66), [1, 3, 3, 1]), (1, 1, (1, 1), (1, 2, 257, 257), [1, 2, 1]),
    (2, 1, (1, 2, 0, 3), (1, 2, 33, 64), [1, 2, 1]), (1, 2, (2, 0, 1, 1), (1, 3, 90, 131), [1, 3, 3, 1])])
def test_upfirdn2d_planar_path_vs_oracle(L, up, down, pad, shape, taps):
    """NCHW drop-in calls large enough for the tiled planar kernel (minor == 1), against the oracle's upfirdn2d"""
    from ganecdotes_b200.stylegan2.op import upfirdn2d
    g = torch.Generator().manual_seed(sum(shape) + up + 2 * down)
    x = torch.randn(*shape, generator=g)
    k = O.make_fir_kernel(taps) * (up ** 2)
    k = k + 0.01 * torch.randn(k.shape, generator=g)           # non-symmetric: the flip matters
    ref = O.upfirdn2d(x, k, up=up, down=down, pad=pad)
    got = upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad)
    assert got.shape == ref.shape
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("cfg", [
    dict(b=2, cin=6, cout=8, hw=(9, 11), k=3, stride=1, padding=1, dilation=1, groups=1),
    dict(b=1, cin=8, cout=12, hw=(16, 16), k=3, stride=2, padding=1, dilation=1, groups=2),
    dict(b=2, cin=4, cout=4, hw=(10, 7), k=(3, 2), stride=(1, 2), padding=(2, 0), dilation=(2, 1), groups=1),
    dict(b=3, cin=16, cout=16, hw=(8, 8), k=1, stride=1, padding=0, dilation=1, groups=1),
])
def test_conv2d_gradfix_matches_torch_to_second_order(L, cfg):
    """conv2d / conv_transpose2d of the conv2d_gradfix drop-in (ref lib/gan/optim/conv2d_gradfix.py) against
    torch.nn.functional (what the reference itself calls on torch >= 1.9): values, first-order gradients and a
    gradient penalty (gradient of a squared gradient norm: second order), plus no_weight_gradients()."""
    import torch.nn.functional as F
    from ganecdotes_b200.stylegan2.op import conv2d_gradfix as G
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator().manual_seed(cfg["cin"] + cfg["cout"])
        k = cfg["k"] if isinstance(cfg["k"], tuple) else (cfg["k"], cfg["k"])
        kw = dict(stride=cfg["stride"], padding=cfg["padding"], dilation=cfg["dilation"], groups=cfg["groups"])

        def mk(*shape):
            return torch.randn(*shape, generator=g).cuda().requires_grad_(True)
        for transpose in (False, True):
            if transpose:
                x = mk(cfg["b"], cfg["cin"], *cfg["hw"])
                w = mk(cfg["cin"], cfg["cout"] // cfg["groups"], *k)
                bias = mk(cfg["cout"])
                s = cfg["stride"] if isinstance(cfg["stride"], tuple) else (cfg["stride"],) * 2
                opad = tuple(1 if si > 1 else 0 for si in s)
                f_ref = lambda x_, w_, b_: F.conv_transpose2d(x_, w_, b_, output_padding=opad, **kw)
                f_got = lambda x_, w_, b_: G.conv_transpose2d(x_, w_, b_, output_padding=opad, **kw)
            else:
                x = mk(cfg["b"], cfg["cin"], *cfg["hw"])
                w = mk(cfg["cout"], cfg["cin"] // cfg["groups"], *k)
                bias = mk(cfg["cout"])
                f_ref = lambda x_, w_, b_: F.conv2d(x_, w_, b_, **kw)
                f_got = lambda x_, w_, b_: G.conv2d(x_, w_, b_, **kw)
            y_ref, y_got = f_ref(x, w, bias), f_got(x, w, bias)
            assert y_got.shape == y_ref.shape
            scale = y_ref.abs().max().item()
            assert (y_got - y_ref).abs().max().item() < 1e-4 * scale
            go = torch.randn(y_ref.shape, generator=g).cuda()
            r1 = torch.autograd.grad(y_ref, (x, w, bias), go, create_graph=True)
            g1 = torch.autograd.grad(y_got, (x, w, bias), go, create_graph=True)
            for a, b_ in zip(g1, r1):
                assert (a - b_).abs().max().item() < 2e-4 * b_.abs().max().item()
            # gradient penalty: d/d(w, x) of |dy/dx|^2  (second order through the conv and through its weight gradient)
            pen_ref = r1[0].pow(2).sum() + r1[1].pow(2).sum()
            pen_got = g1[0].pow(2).sum() + g1[1].pow(2).sum()
            r2 = torch.autograd.grad(pen_ref, (x, w))
            g2 = torch.autograd.grad(pen_got, (x, w))
            for a, b_ in zip(g2, r2):
                assert (a - b_).abs().max().item() < 5e-4 * b_.abs().max().item()
            with G.no_weight_gradients():
                y = f_got(x, w, bias)
                gx, = torch.autograd.grad(y, (x,), go, retain_graph=True)
                assert (gx - r1[0].detach()).abs().max().item() < 2e-4 * r1[0].abs().max().item()
                assert torch.autograd.grad(y, (w,), go, allow_unused=True)[0] is None
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_native_ops_half_and_double(L):
    """float16 / float64 instantiations of upfirdn2d and fused_bias_act (the reference's
    AT_DISPATCH_FLOATING_TYPES_AND_HALF) - against the reference's own CUDA kernels when oracle/_ref is built, and
    against the fp64 oracle formulas in any case"""
    from ganecdotes_b200.stylegan2.op import upfirdn2d, fused_leaky_relu
    from oracle import build_ref
    up, fb = build_ref.load("ref_upfirdn2d"), build_ref.load("ref_fused_bias_act")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 20, 22, generator=g, dtype=torch.float64)
    k = O.make_fir_kernel([1, 3, 3, 1]).double() * 4
    bias = torch.randn(3, generator=g, dtype=torch.float64)
    for dt, tol in ((torch.float64, 1e-12), (torch.float16, 4e-3)):
        for upf, down, pad in ((2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (2, 1))):
            ref = O.upfirdn2d(x, k, up=upf, down=down, pad=pad)
            got = upfirdn2d(x.cuda().to(dt), k.cuda().to(dt), up=upf, down=down, pad=pad)
            assert got.dtype == dt
            assert (got.double().cpu() - ref).abs().max().item() < tol * ref.abs().max().item()
            if up is not None:
                xr = x.cuda().to(dt).reshape(-1, 20, 22, 1).contiguous()
                r2 = up.upfirdn2d(xr, k.cuda().to(dt), upf, upf, down, down, pad[0], pad[1], pad[0], pad[1])
                # (the reference's own float64 instantiation is only float32-accurate - 1.3e-7 from the fp64 formula,
                # which this kernel matches to 1e-12 - so the comparison with it is at float32 accuracy)
                assert (got.reshape(r2.shape).double() - r2.double()).abs().max().item() <= max(2 * tol, 1e-6) * ref.abs().max().item()
        ref = O.fused_leaky_relu(x, bias)
        got = fused_leaky_relu(x.cuda().to(dt), bias.cuda().to(dt))
        assert got.dtype == dt
        # alpha and scale cross the ABI as C floats, like the reference's op (fused_bias_act.cpp:18-24): sqrt(2) is
        # float32-accurate even in the float64 instantiation
        assert (got.double().cpu() - ref).abs().max().item() < max(tol, 1e-7) * ref.abs().max().item()
        if fb is not None:
            r2 = fb.fused_bias_act(x.cuda().to(dt), bias.cuda().to(dt), torch.empty(0, device="cuda", dtype=dt), 3, 0, 0.2,
                                   2 ** 0.5)
            assert (got.double() - r2.double()).abs().max().item() <= max(2 * tol, 1e-6) * ref.abs().max().item()
