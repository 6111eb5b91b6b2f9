"""ctypes binding of the C ABI (include/ganecdotes_b200.h) + thin torch-tensor wrappers.

There is no CPU fallback: if the shared library is missing, or the device is not a
CUDA sm_100 device, every op raises RuntimeError.
"""
import ctypes as C
import os
import threading

import torch

from . import build as _build

_LIB = None
_LOCK = threading.Lock()
GX_MAX_LEVELS = 16

# number of CUDA kernels launched through this module (claimed in bench.py's gpu_launches)
launch_count = 0
# separable blur kernel for separable filters (GX_BLUR_SEP=0 / 1 overrides; A/B timing and tests)
BLUR_SEP_DEFAULT = "1"


class GxError(RuntimeError):
    pass


class gx_conv_desc(C.Structure):
    _fields_ = [
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
        ("batch", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
        ("cin_ld", C.c_int), ("upsample", C.c_int), ("passes", C.c_int),
        ("demod", C.c_void_p), ("noise", C.c_void_p), ("noise_batch_stride", C.c_longlong),
        ("noise_strength", C.c_void_p), ("bias", C.c_void_p), ("act", C.c_int),
        ("out", C.c_void_p), ("next_style", C.c_void_p), ("next_hi", C.c_void_p), ("next_lo", C.c_void_p),
        ("next_ld", C.c_int), ("block_n", C.c_int), ("stages", C.c_int), ("cluster_pair", C.c_int),
        ("dilation", C.c_int),
    ]


class gx_gemm_desc(C.Structure):
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("b_hi", C.c_void_p), ("b_lo", C.c_void_p),
        ("lda", C.c_longlong), ("ldb", C.c_longlong), ("a_mn_major", C.c_int), ("b_mn_major", C.c_int),
        ("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("passes", C.c_int),
        ("c", C.c_void_p), ("ldc", C.c_longlong), ("bias", C.c_void_p), ("split_k", C.c_int),
        ("accumulate", C.c_int), ("force_m128", C.c_int), ("colexp_sum", C.c_void_p), ("colexp_scale", C.c_float),
        ("block_n", C.c_int), ("stages", C.c_int), ("cluster_pair", C.c_int), ("ab_f16", C.c_int),
    ]


class gx_gather_desc(C.Structure):
    _fields_ = [
        ("nlevels", C.c_int),
        ("feat", C.c_void_p * GX_MAX_LEVELS),
        ("h", C.c_int * GX_MAX_LEVELS), ("w", C.c_int * GX_MAX_LEVELS), ("c", C.c_int * GX_MAX_LEVELS),
        ("out_h", C.c_int), ("out_w", C.c_int), ("hlen", C.c_int),
        ("row_img", C.c_void_p), ("row_src", C.c_void_p), ("nrows", C.c_longlong),
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("a_f32", C.c_void_p), ("ld", C.c_longlong),
        ("row_norm", C.c_void_p),
    ]


GX_MAX_PEERS = 16


class gx_ll_desc(C.Structure):
    _fields_ = [("peers", C.c_void_p * GX_MAX_PEERS), ("world", C.c_int), ("rank", C.c_int),
                ("block_words", C.c_longlong), ("seq", C.c_uint), ("err", C.c_void_p)]


_I, _LL, _F, _P = C.c_int, C.c_longlong, C.c_float, C.c_void_p
_LLD = C.POINTER(gx_ll_desc)

_SIGNATURES = {
    "gx_version": ([], _I),
    "gx_abi_sizeof": ([_I], _I),
    "gx_last_cuda_error": ([], _I),
    "gx_error_string": ([_I], C.c_char_p),
    "gx_device_ok": ([], _I),
    "gx_set_sm_budget": ([_I, _I], _I),
    "gx_upfirdn2d": ([_P, _P, _P] + [_I] * 14 + [_P], _I),
    "gx_fused_bias_act": ([_P, _P, _P, _P, _LL, _I, _I, _I, _I, _F, _F, _P], _I),
    "gx_upfirdn2d_t": ([_I, _P, _P, _P] + [_I] * 14 + [_P], _I),
    "gx_fused_bias_act_t": ([_I, _P, _P, _P, _P, _LL, _I, _I, _I, _I, _F, _F, _P], _I),
    "gx_pixel_norm": ([_P, _P, _I, _I, _P], _I),
    "gx_equal_linear": ([_P, _LL, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P], _I),
    "gx_truncate": ([_P, _P, _P, _LL, _I, _F, _P], _I),
    "gx_view_wplus": ([_P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _P, _P], _I),
    "gx_pixel_segments_scratch": ([_LL], _I),
    "gx_pixel_segments": ([_P, _P, _I, _LL, _I, _LL, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    "gx_colsum": ([_P, _I, _I, _F, _I, _P, _P], _I),
    "gx_modconv_prepare": ([_P, _F, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "gx_modconv_demod": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "gx_modulate_split": ([_P, _LL, _P, _P, _P, _I, _LL, _I, _I, _P], _I),
    "gx_modconv": ([C.POINTER(gx_conv_desc), _P], _I),
    "gx_modconv_small": ([_P, _P, _P, _P, _P, _LL, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "gx_modconv_small_supported": ([_I, _I], _I),
    "gx_blur_noise_bias_act": ([_P, _P, _I, _I, _I, _I, _P, _LL, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "gx_blur_sep_noise_bias_act": ([_P, _P, _P, _I, _I, _I, _P, _LL, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "gx_torgb": ([_P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _P], _I),
    "gx_gemm": ([C.POINTER(gx_gemm_desc), _P], _I),
    "gx_gemm_check": ([C.POINTER(gx_gemm_desc), _P], _I),
    "gx_split_planes": ([_P, _LL, _P, _P, _LL, _LL, _I, _LL, _P], _I),
    "gx_gather_rows": ([C.POINTER(gx_gather_desc), _P], _I),
    "gx_l2norm_split": ([_P, _P, _P, _P, _P, _P, _LL, _I, _P], _I),
    "gx_round_f16": ([_P, _LL, _P, _LL, _LL, _P], _I),
    "gx_l2norm_bwd_split": ([_P, _P, _P, _P, _P, _P, _P, _LL, _I, _P], _I),
    "gx_segment_sum_rows": ([_P, _I, _P, _P, _P, _P, _P, _LL, _I, _P], _I),
    "gx_upsample_sum": ([_I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P], _I),
    "gx_pool1d_bilinear": ([_P, _LL, _I, _I, _LL, _P, _P], _I),
    "gx_pool_sum": ([_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "gx_tap_sum": ([_P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _P], _I),
    "gx_tap_spread": ([_P, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "gx_normalize_rows": ([_P, _LL, _I, _P], _I),
    "gx_sinkhorn_max_parts": ([], _I),
    "gx_peer_alloc": ([_LL, C.POINTER(_P)], _I),
    "gx_peer_free": ([_P], _I),
    "gx_peer_export": ([_P, _P], _I),
    "gx_peer_open": ([_P, C.POINTER(_P)], _I),
    "gx_peer_close": ([_P], _I),
    "gx_sinkhorn_pass": ([_P, _LL, _I, _LL, _F, _I, _P, _LLD, _P, _P, _LL, _I, _P, C.POINTER(_I), _P], _I),
    "gx_sinkhorn_pass_cached": ([_P, _LL, _I, _LL, _F, _P, _LLD, _P, _P, _LL, _I, _P, C.POINTER(_I), _P, _LL, _P, _I,
                                 _P], _I),
    "gx_sinkhorn_reduce": ([_P, _I, _I, _P, _P], _I),
    "gx_sinkhorn_reduce_send": ([_P, _I, _I, _LLD, _P, _P], _I),
    "gx_ll_recv_sum": ([_LLD, _I, _P, _P], _I),
    "gx_sinkhorn_log_a": ([_P, _LLD, _P, _I, _P, _P], _I),
    "gx_sinkhorn_q": ([_P, _LL, _I, _LL, _F, _P, _P, _P], _I),
    "gx_loss_max_parts": ([], _I),
    "gx_swav_loss": ([_P, _P, _LL, _I, _LL, _F, _F, _P, _P, _F, _P, _P, C.POINTER(_I), _P, _P, _P, _P, _LL, _P, _P,
                      _P], _I),
    "gx_im2col": ([_P] + [_I] * 14 + [_LL, _P, _P], _I),
    "gx_col2im": ([_P] + [_I] * 14 + [_LL, _P, _P], _I),
    "gx_recip": ([_P, _LL, _F, _I, _P, _P], _I),
    "gx_bn_stats": ([_P, _LL, _P, _I, _I, _F, _P, _P, _P, _P, _F, _P], _I),
    "gx_bn_act_apply": ([_P, _LL, _P, _LL, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P], _I),
    "gx_bn_act_bwd": ([_P, _P, _LL, _P, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P], _I),
    "gx_simclr_loss": ([_P, _I, _I, _F, _P, _P, _P], _I),
    "gx_larc_scratch_floats": ([], _I),
    "gx_larc_sgd": ([_P, _P, _P, _LL, _F, _F, _F, _F, _F, _I, _P, _P], _I),
    "gx_argmax_rows": ([_P, _LL, _I, _LL, _P, _P], _I),
    "gx_kmeans_assign": ([_P, _I, _P, _I, _LL, _P, _I, _P, _P, _P], _I),
    "gx_argmin_affine": ([_P, _LL, _I, _LL, _P, _F, _P, _P], _I),
    "gx_kmeans_frag_bytes": ([_I, _I], _LL),
    "gx_kmeans_center_frags": ([_P, _I, _I, _P, _P], _I),
    "gx_kmeans_assign_mma": ([_P, _I, _P, _I, _LL, _P, _P, _I, _P, _P], _I),
    "gx_onehot_nearest": ([_P, _I, _I, _I, _I, _I, _I, _P, _LL, _F, _F, _P], _I),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())


def lib_path() -> str:
    return _build.LIB_PATH


GX_ABI_VERSION = 203        # include/ganecdotes_b200.h


def load(require_device: bool = True):
    """Load the CUDA library, rebuilding it first when it is missing or older than its sources (nvcc
    permitting), then check the ABI handshake: version + sizeof of every descriptor struct."""
    global _LIB
    with _LOCK:
        if _LIB is None:
            path = _build.LIB_PATH
            try:
                _build.build()          # returns at once when the .so is newer than every source / header
            except Exception as e:
                if not os.path.exists(path):
                    raise GxError(f"ganecdotes_b200: CUDA library missing and could not be built: {e}")
                # no nvcc on this box: a shipped .so is used as is, the handshake below still applies
            lib = C.CDLL(path)
            for name, (argt, rest) in _SIGNATURES.items():
                try:
                    fn = getattr(lib, name)
                except AttributeError:
                    raise GxError(f"ganecdotes_b200: {path} is stale (no symbol {name}); rebuild with "
                                  f"`python -m ganecdotes_b200.build`")
                fn.argtypes = argt
                fn.restype = rest
            if lib.gx_version() != GX_ABI_VERSION:
                raise GxError(f"ganecdotes_b200: library ABI {lib.gx_version()} != binding ABI {GX_ABI_VERSION}; "
                              f"rebuild with `python -m ganecdotes_b200.build`")
            for which, st in enumerate((gx_conv_desc, gx_gemm_desc, gx_gather_desc, gx_ll_desc)):
                if lib.gx_abi_sizeof(which) != C.sizeof(st):
                    raise GxError(f"ganecdotes_b200: sizeof({st.__name__}) is {lib.gx_abi_sizeof(which)} in the library "
                                  f"and {C.sizeof(st)} in the binding")
            _LIB = lib
    if require_device:
        if not torch.cuda.is_available():
            raise GxError("ganecdotes_b200 has no CPU fallback: a CUDA sm_100 (B200) device is required")
        if not getattr(load, "_dev_ok", False):
            if _LIB.gx_device_ok() != 1:
                raise GxError("ganecdotes_b200 kernels are built for sm_100a only; current device is not sm_100")
            load._dev_ok = True
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        lib = _LIB
        msg = lib.gx_error_string(rc).decode()
        extra = ""
        if rc == -2:
            extra = f" (cuda error {lib.gx_last_cuda_error()})"
        raise GxError(f"{what}: {msg}{extra}")


def set_sm_budget(umma_ctas=0, stream_ctas=0):
    """CTA budgets of the persistent kernel families (0 = all SMs)."""
    _check(load().gx_set_sm_budget(int(umma_ctas), int(stream_ctas)), "gx_set_sm_budget")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise GxError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise GxError(f"{name} must be float32")
    if not t.is_contiguous():
        raise GxError(f"{name} must be contiguous")
    return t


def _count(n=1):
    global launch_count
    launch_count += n


# Optional per-kernel CUDA-event timing (used by bench.py for the live roofline numbers).
# event_log: None (off) or a list of (name, start_event, end_event, work) tuples.
event_log = None


class timed:
    """with timed("name", work): launch...  - records CUDA events on the launching stream."""

    def __init__(self, name, work=0.0, nbytes=0.0):
        # work: algorithmic FLOPs (tensor-bound stages) or bytes (HBM-bound stages); nbytes: for a tensor stage, the
        # bytes it must move (operands once + output once) - bench.py reports the stage against whichever roofline
        # bounds it (a short-K GEMM is bound by its output, not by the tensor pipe)
        self.name, self.work, self.nbytes = name, work, nbytes

    def __enter__(self):
        if event_log is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if event_log is not None:
            self.e1.record()
            event_log.append((self.name, self.e0, self.e1, self.work, self.nbytes))
        return False


# ----------------------------------------------------------------------------------------
# op wrappers (raw): callers pass contiguous CUDA tensors of the right dtype
# ----------------------------------------------------------------------------------------

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


def _typed(t, name, dtype):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise GxError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")
    return t


def upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """x4: [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]; float32, float16 or float64 (taps in the
    input's type, like the reference's dispatch)"""
    lib = load()
    if x4.dtype not in _DTYPE_CODE:
        raise GxError("upfirdn2d: float32, float16 or float64 input")
    _typed(x4, "input", x4.dtype), _typed(kernel, "kernel", x4.dtype)
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    out_h = (in_h * up_y + py0 + py1 - kh + down_y) // down_y
    out_w = (in_w * up_x + px0 + px1 - kw + down_x) // down_x
    if out_h <= 0 or out_w <= 0:
        raise GxError("upfirdn2d: empty output")
    out = torch.empty((major, out_h, out_w, minor), dtype=x4.dtype, device=x4.device)
    if out.numel():
        _check(lib.gx_upfirdn2d_t(_DTYPE_CODE[x4.dtype], _ptr(x4), _ptr(kernel), _ptr(out), major, in_h, in_w, minor, kh,
                                  kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1, _stream()), "gx_upfirdn2d")
        _count()
    return out


def fused_bias_act_raw(x, bias, refer, act, grad, alpha, scale, out=None):
    """float32, float16 or float64 (bias / refer in the input's type); `out`: a contiguous tensor of the input's
    type and size to write into (default: a new one)"""
    lib = load()
    if x.dtype not in _DTYPE_CODE:
        raise GxError("fused_bias_act: float32, float16 or float64 input")
    _typed(x, "input", x.dtype)
    if out is None:
        out = torch.empty_like(x)
    else:
        _typed(out, "out", x.dtype)
        if out.numel() != x.numel():
            raise GxError("fused_bias_act: out must have the input's size")
    if x.numel() == 0:
        return out
    if bias is not None:
        _typed(bias, "bias", x.dtype)
        step_b = 1
        for d in x.shape[2:]:
            step_b *= d
        size_b = x.shape[1]
        if bias.numel() != size_b:
            raise GxError("fused_bias_act: bias must have input.shape[1] elements")
    else:
        step_b = size_b = 1
    if refer is not None:
        _typed(refer, "refer", x.dtype)
    _check(lib.gx_fused_bias_act_t(_DTYPE_CODE[x.dtype], _ptr(x), _ptr(bias), _ptr(refer), _ptr(out), x.numel(), step_b,
                                   size_b, act, grad, float(alpha), float(scale), _stream()), "gx_fused_bias_act")
    _count()
    return out


def pixel_norm(x):
    lib = load()
    _f32(x, "x")
    y = torch.empty_like(x)
    _check(lib.gx_pixel_norm(_ptr(x), _ptr(y), x.shape[0], x.shape[1], _stream()), "gx_pixel_norm")
    _count()
    return y


def equal_linear(x, w, b, w_scale, b_scale, act):
    """x [n, in_dim] fp32 with unit inner stride (rows may be strided: a row of W+ is read in place)"""
    lib = load()
    _f32(w, "w"), _f32(b, "b")
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise GxError("equal_linear: x must be a 2-D float32 CUDA tensor with unit inner stride")
    n, in_dim = x.shape
    ldx = x.stride(0) if n > 1 else in_dim
    out_dim = w.shape[0]
    y = torch.empty((n, out_dim), dtype=torch.float32, device=x.device)
    # grid.y limit: chunk rows
    step = 65535 * 8
    for r0 in range(0, n, step):
        xs = x[r0:r0 + step]
        ys = y[r0:r0 + step]
        _check(lib.gx_equal_linear(_ptr(xs), ldx, _ptr(w), _ptr(b), _ptr(ys), xs.shape[0], in_dim, out_dim,
                                   float(w_scale), float(b_scale), int(act), _stream()), "gx_equal_linear")
        _count()
    return y


def view_wplus(w, noise_w, layer_no, sigma, mean, psi, n_latent):
    """W+ [rows, n_latent, D] of both perturbed views (gx_view_wplus): w [b, D], noise_w [2*rows, D],
    layer_no int32 [rows], sigma fp32 [rows] (device tensors)"""
    lib = load()
    _f32(w, "w"), _f32(noise_w, "noise_w"), _f32(sigma, "sigma"), _f32(mean, "mean")
    b, d = w.shape
    rows = layer_no.numel()
    if layer_no.dtype != torch.int32 or noise_w.shape[0] != 2 * rows or sigma.numel() != rows or rows % b:
        raise GxError("view_wplus: layer_no int32 [rows], sigma [rows], noise_w [2*rows, D], rows a multiple of b")
    out = torch.empty((rows, n_latent, d), dtype=torch.float32, device=w.device)
    _check(lib.gx_view_wplus(_ptr(w), _ptr(noise_w), _ptr(layer_no), _ptr(sigma), _ptr(mean), float(psi), b, rows,
                             int(n_latent), d, _ptr(out), _stream()), "gx_view_wplus")
    _count()
    return out


def pixel_segments(row_src, row_img, hw, npix, patches_per_group=0, img_group_stride=0):
    """(ridx [P, bn], order [P*bn], seg_off [npix+1]) - gx_pixel_segments; int32 device tensors.  Several views
    stacked along the patch dim: view v = patch // patches_per_group, its images start at v * img_group_stride."""
    lib = load()
    patches, bn = row_src.shape
    dev = row_src.device
    if row_src.dtype != torch.int32 or row_img.dtype != torch.int32 or not row_src.is_contiguous():
        raise GxError("pixel_segments: int32 contiguous row_src / row_img")
    ridx = torch.empty((patches, bn), dtype=torch.int32, device=dev)
    order = torch.empty((patches * bn,), dtype=torch.int32, device=dev)
    seg_off = torch.empty((npix + 1,), dtype=torch.int32, device=dev)
    counts = torch.empty((npix,), dtype=torch.int32, device=dev)
    scratch = torch.empty((lib.gx_pixel_segments_scratch(npix),), dtype=torch.int32, device=dev)
    _check(lib.gx_pixel_segments(_ptr(row_src), _ptr(row_img), patches, bn, int(hw), int(npix), int(patches_per_group),
                                 int(img_group_stride), _ptr(ridx), _ptr(counts), _ptr(scratch), _ptr(seg_off),
                                 _ptr(order), _stream()),
           "gx_pixel_segments")
    _count(6)
    return ridx, order, seg_off


def colsum(parts, nparts, k, out, scale=1.0, accumulate=False):
    """out[k] (+)= scale * sum_p parts[p, k]  (gx_colsum)"""
    _check(load().gx_colsum(_ptr(parts), int(nparts), int(k), float(scale), int(bool(accumulate)), _ptr(out),
                            _stream()), "gx_colsum")
    _count()
    return out


def truncate(w, mean, psi):
    lib = load()
    _f32(w, "w"), _f32(mean, "mean")
    out = torch.empty_like(w)
    dim = w.shape[-1]
    _check(lib.gx_truncate(_ptr(w), _ptr(mean), _ptr(out), w.numel() // dim, dim, float(psi), _stream()),
           "gx_truncate")
    _count()
    return out


def pad64(c):
    """channels per pixel of the bf16 operand planes: multiple of the 64-element K block"""
    return (c + 63) // 64 * 64


def _planes(shape, dev, padded, want_lo=True):
    alloc = torch.zeros if padded else torch.empty
    hi = alloc(shape, dtype=torch.bfloat16, device=dev)
    lo = alloc(shape, dtype=torch.bfloat16, device=dev) if want_lo else None
    return hi, lo


def modconv_prepare(weight, scale, want_lo=True):
    """weight [cout,cin,k,k] -> (w_hi, w_lo) [cout, k*k*pad64(cin)] bf16, wsq [cout,cin]"""
    lib = load()
    _f32(weight, "weight")
    cout, cin, k, _ = weight.shape
    cin_ld = pad64(cin)
    w_hi = torch.empty((cout, k * k * cin_ld), dtype=torch.bfloat16, device=weight.device)
    w_lo = torch.empty_like(w_hi) if want_lo else None
    wsq = torch.empty((cout, cin), dtype=torch.float32, device=weight.device)
    _check(lib.gx_modconv_prepare(_ptr(weight), float(scale), _ptr(w_hi), _ptr(w_lo), _ptr(wsq), cout, cin, cin_ld,
                                  k, _stream()), "gx_modconv_prepare")
    _count()
    return w_hi, w_lo, wsq


def modconv_demod(wsq, s):
    lib = load()
    _f32(wsq, "wsq"), _f32(s, "s")
    cout, cin = wsq.shape
    b = s.shape[0]
    d = torch.empty((b, cout), dtype=torch.float32, device=s.device)
    _check(lib.gx_modconv_demod(_ptr(wsq), _ptr(s), _ptr(d), b, cin, cout, _stream()), "gx_modconv_demod")
    _count()
    return d


def modulate_split(x_nhwc, s, batch, want_lo=True):
    """x_nhwc: [B or 1, H, W, C] fp32; s [batch, C] -> hi, lo [batch,H,W,pad64(C)] bf16"""
    lib = load()
    _f32(x_nhwc, "x"), _f32(s, "s")
    xb, h, w, c = x_nhwc.shape
    stride = 0 if (xb == 1 and batch > 1) else h * w * c
    c_ld = pad64(c)
    hi, lo = _planes((batch, h, w, c_ld), x_nhwc.device, c_ld != c, want_lo)
    _check(lib.gx_modulate_split(_ptr(x_nhwc), stride, _ptr(s), _ptr(hi), _ptr(lo), batch, h * w, c, c_ld,
                                 _stream()), "gx_modulate_split")
    _count()
    return hi, lo


def modconv(x_hi, x_lo, w_hi, w_lo, cout, upsample, passes, demod=None, noise=None, noise_strength=None, bias=None,
            act=0, next_style=None, want_next_lo=True, block_n=0, stages=0, tag="modconv", cin_true=None, dilation=1, pair=True):
    """Implicit-GEMM modulated conv.  x_*: [B,H,W,pad64(Cin)] bf16 planes.
    Returns (out fp32 NHWC [B,Ho,Wo,cout], next_hi, next_lo [B,Ho,Wo,pad64(cout)])."""
    lib = load()
    b, h, w, cin_ld = x_hi.shape
    cin = cin_ld
    ho, wo = (2 * h + 1, 2 * w + 1) if upsample else (h, w)
    dev = x_hi.device
    out = torch.empty((b, ho, wo, cout), dtype=torch.float32, device=dev)
    next_hi = next_lo = None
    next_ld = pad64(cout)
    if next_style is not None:
        next_hi, next_lo = _planes((b, ho, wo, next_ld), dev, next_ld != cout, want_next_lo)
    d = gx_conv_desc()
    d.x_hi, d.x_lo, d.w_hi, d.w_lo = _ptr(x_hi), _ptr(x_lo), _ptr(w_hi), _ptr(w_lo)
    d.batch, d.h, d.w, d.cin, d.cout, d.cin_ld = b, h, w, cin, cout, cin_ld
    d.upsample, d.passes = int(bool(upsample)), passes
    d.demod = _ptr(demod)
    d.noise = _ptr(noise)
    if noise is not None:
        d.noise_batch_stride = 0 if noise.shape[0] == 1 else ho * wo
        d.noise_strength = _ptr(noise_strength)
    d.bias = _ptr(bias)
    d.act = int(act)
    d.out = _ptr(out)
    d.next_style, d.next_hi, d.next_lo = _ptr(next_style), _ptr(next_hi), _ptr(next_lo)
    d.next_ld = next_ld
    d.block_n, d.stages = block_n, stages
    d.dilation = int(dilation)
    d.cluster_pair = int(pair)
    npl = 2 if passes == 3 else 1
    conv_bytes = (2.0 * npl * b * h * w * cin_ld + 4.0 * b * ho * wo * cout
                  + (2.0 * (2 if next_lo is not None else 1) * b * ho * wo * next_ld if next_hi is not None else 0.0)
                  + 2.0 * npl * 9 * cin_ld * cout)          # input planes once + outputs + weights
    with timed(tag + ("_up" if upsample else ""), 2.0 * b * h * w * 9 * (cin_true or cin) * cout, conv_bytes):
        _check(lib.gx_modconv(C.byref(d), _stream()), "gx_modconv")
    _count()
    return out, next_hi, next_lo


def modconv_small_supported(cin, cout):
    return bool(load().gx_modconv_small_supported(int(cin), int(cout)))


def modconv_small_weights(weight, scale):
    """[cout, cin, 3, 3] conv weight -> fp32 [9, cin, cout] with the equalised-lr scale folded in (operand of
    `modconv_small`; prepared once per weight version, like the bf16 planes of the tcgen05 path)"""
    return (weight.detach().float() * float(scale)).permute(2, 3, 1, 0).reshape(9, weight.shape[1],
                                                                                weight.shape[0]).contiguous()


def modconv_small(x_nhwc, style, w9, demod=None, noise=None, noise_strength=None, bias=None, act=0, next_style=None,
                  want_next_lo=True, tag="modconv_small"):
    """Direct fp32 modulated 3x3 conv for few-channel layers.  x_nhwc fp32 [B,H,W,cin] (unmodulated), style [B,cin],
    w9 = modconv_small_weights(...).  Returns (out fp32 NHWC [B,H,W,cout], next_hi, next_lo [B,H,W,pad64(cout)])."""
    lib = load()
    _f32(x_nhwc, "x"), _f32(style, "style"), _f32(w9, "w9")
    b, h, w, cin = x_nhwc.shape
    cout = w9.shape[2]
    if w9.shape[0] != 9 or w9.shape[1] != cin or style.shape != (b, cin):
        raise GxError("modconv_small: w9 must be [9, cin, cout] and style [B, cin]")
    dev = x_nhwc.device
    out = torch.empty((b, h, w, cout), dtype=torch.float32, device=dev)
    next_hi = next_lo = None
    next_ld = pad64(cout)
    if next_style is not None:
        next_hi, next_lo = _planes((b, h, w, next_ld), dev, next_ld != cout, want_next_lo)
    nbs = 0
    if noise is not None:
        nbs = 0 if noise.shape[0] == 1 else h * w
    nbytes = 4.0 * b * h * w * (cin + cout) + (2.0 * (2 if next_lo is not None else 1) * b * h * w * next_ld
                                                if next_hi is not None else 0.0)
    with timed(tag, 2.0 * b * h * w * 9 * cin * cout, nbytes):
        _check(lib.gx_modconv_small(_ptr(x_nhwc), _ptr(style), _ptr(w9), _ptr(demod), _ptr(noise), nbs,
                                    _ptr(noise_strength), _ptr(bias), int(act), _ptr(out), _ptr(next_style),
                                    _ptr(next_hi), _ptr(next_lo), next_ld, b, h, w, cin, cout, _stream()),
               "gx_modconv_small")
    _count()
    return out, next_hi, next_lo


def separable_factors(fir):
    """(fir_x, fir_y) device vectors with fir = outer(fir_y, fir_x) to fp32 rounding, or None.  One host read of the
    4x4 filter; callers cache the result (Blur.separable)."""
    k = fir.detach().double().cpu()
    if k.dim() != 2 or k.abs().max() == 0:
        return None
    i, j = divmod(int(k.abs().argmax()), k.shape[1])
    piv = k[i, j]
    root = piv.abs().sqrt()
    fy = k[:, j] / piv * root * (1.0 if piv > 0 else -1.0)
    fx = k[i, :] / root
    if (torch.outer(fy, fx) - k).abs().max() > 1e-7 * k.abs().max():
        return None
    return fx.float().to(fir.device).contiguous(), fy.float().to(fir.device).contiguous()


def blur_noise_bias_act(x_nhwc, fir, pad0, pad1, noise, noise_strength, bias, act, next_style, want_next_lo=True,
                        sep=None):
    """`sep=(fir_x, fir_y)` (separable_factors(fir)): the separable kernel, 4x fewer multiply-adds."""
    lib = load()
    _f32(x_nhwc, "x"), _f32(fir, "fir")
    b, hi, wi, c = x_nhwc.shape
    kh, kw = fir.shape
    ho, wo = hi + pad0 + pad1 - kh + 1, wi + pad0 + pad1 - kw + 1
    dev = x_nhwc.device
    out = torch.empty((b, ho, wo, c), dtype=torch.float32, device=dev)
    next_hi = next_lo = None
    next_ld = pad64(c)
    if next_style is not None:
        next_hi, next_lo = _planes((b, ho, wo, next_ld), dev, next_ld != c, want_next_lo)
    nbs = 0
    if noise is not None:
        nbs = 0 if noise.shape[0] == 1 else ho * wo
    nbytes = 4.0 * b * c * (hi * wi + ho * wo) + (2.0 * b * c * ho * wo * (2 if next_lo is not None else 1)
                                                   if next_hi is not None else 0.0)
    if sep is not None and kh == 4 and kw == 4 and os.environ.get("GX_BLUR_SEP", BLUR_SEP_DEFAULT) != "0":
        with timed("blur_noise_bias_act", nbytes):
            _check(lib.gx_blur_sep_noise_bias_act(_ptr(x_nhwc), _ptr(sep[0]), _ptr(sep[1]), 4, pad0, pad1, _ptr(noise),
                                                  nbs, _ptr(noise_strength), _ptr(bias), int(act), _ptr(out),
                                                  _ptr(next_style), _ptr(next_hi), _ptr(next_lo), next_ld, b, hi, wi,
                                                  c, _stream()), "gx_blur_sep_noise_bias_act")
        _count()
        return out, next_hi, next_lo
    with timed("blur_noise_bias_act", nbytes):
        _check(lib.gx_blur_noise_bias_act(_ptr(x_nhwc), _ptr(fir), kh, kw, pad0, pad1, _ptr(noise), nbs,
                                          _ptr(noise_strength), _ptr(bias), int(act), _ptr(out), _ptr(next_style),
                                          _ptr(next_hi), _ptr(next_lo), next_ld, b, hi, wi, c, _stream()),
               "gx_blur_noise_bias_act")
    _count()
    return out, next_hi, next_lo


def torgb(x_nhwc, w3c, w_scale, s, bias3, skip):
    lib = load()
    b, h, w, c = x_nhwc.shape
    out = torch.empty((b, 3, h, w), dtype=torch.float32, device=x_nhwc.device)
    _check(lib.gx_torgb(_ptr(x_nhwc), _ptr(w3c), float(w_scale), _ptr(s), _ptr(bias3), _ptr(skip), _ptr(out), b,
                        h * w, c, _stream()), "gx_torgb")
    _count()
    return out


def split_planes(x, transpose=False, want_lo=True, out=None):
    """fp32 [rows, cols] -> bf16 (hi, lo) planes, optionally transposed.  `out=(hi, lo)`: write into
    (column-slice) views of wider planes instead of allocating."""
    lib = load()
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, cols = x.shape
    if out is not None:
        assert not transpose
        hi, lo = out
        assert hi.shape == (rows, cols) and hi.stride(1) == 1 and (lo is None or lo.stride() == hi.stride())
    else:
        shape = (cols, rows) if transpose else (rows, cols)
        hi = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty_like(hi) if want_lo else None
    with timed("split_planes", float(rows) * cols * (4 + (4 if lo is not None else 2))):
        _check(lib.gx_split_planes(_ptr(x), x.stride(0), _ptr(hi), _ptr(lo), rows, cols, int(transpose),
                                   0 if transpose else hi.stride(0), _stream()), "gx_split_planes")
    _count()
    return hi, lo


def round_f16(x):
    """fp32 [rows, cols] -> one fp16 plane"""
    lib = load()
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _check(lib.gx_round_f16(_ptr(x), x.stride(0), _ptr(out), x.shape[0], x.shape[1], _stream()), "gx_round_f16")
    _count()
    return out


def gemm(a_hi, a_lo, b_hi, b_lo, m, n, k, passes, out=None, bias=None, a_mn=False, b_mn=False, split_k=1,
         block_n=0, stages=0, check=False, accumulate=False, tag="gemm", force_m128=False, colexp=None, pair=False):
    """C[m,n] = A * B^T.  Planes are 2-D bf16 tensors: A is [m,k] (or [k,m] if a_mn), B is [n,k] (or [k,n]).
    fp16 planes (both operands, passes == 1) are recognised by dtype."""
    lib = load()
    dev = a_hi.device
    f16 = a_hi.dtype == torch.float16
    if f16 and (b_hi.dtype != torch.float16 or passes != 1):
        raise ValueError("fp16 operands: both planes fp16 and passes == 1")
    if out is None:
        out = (torch.zeros if split_k > 1 else torch.empty)((m, n), dtype=torch.float32, device=dev)
    d = gx_gemm_desc()
    d.a_hi, d.a_lo, d.b_hi, d.b_lo = _ptr(a_hi), _ptr(a_lo), _ptr(b_hi), _ptr(b_lo)
    d.lda, d.ldb = a_hi.stride(0), b_hi.stride(0)
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.m, d.n, d.k, d.passes = m, n, k, passes
    d.c, d.ldc = _ptr(out), out.stride(0)
    d.bias = _ptr(bias)
    d.split_k, d.block_n, d.stages = split_k, block_n, stages
    d.accumulate = int(accumulate)
    d.force_m128 = int(force_m128)
    d.ab_f16 = int(f16)
    d.cluster_pair = int(pair)
    if colexp is not None:      # (zeroed fp32 [n] tensor, scale in the log2 domain)
        d.colexp_sum, d.colexp_scale = _ptr(colexp[0]), float(colexp[1])
    fn = lib.gx_gemm_check if check else lib.gx_gemm
    with timed(tag, 2.0 * m * n * k, (2.0 if passes == 1 else 4.0) * (float(m) * k + float(n) * k) + 4.0 * m * n):
        _check(fn(C.byref(d), _stream()), "gx_gemm")
    _count()
    return out


def gather_rows(feats_nhwc, out_h, out_w, hlen, row_img, row_src, nrows, ld=None, want_lo=True, want_f32=False,
                want_planes=True, want_norm=False):
    """feats_nhwc: list of fp32 [nimg,h,w,c] tensors -> A planes [nrows, ld] bf16"""
    lib = load()
    dev = feats_nhwc[0].device
    ld = ld or hlen
    a_hi = torch.empty((nrows, ld), dtype=torch.bfloat16, device=dev) if want_planes else None
    a_lo = torch.empty_like(a_hi) if (want_lo and want_planes) else None
    a_f = torch.empty((nrows, ld), dtype=torch.float32, device=dev) if want_f32 else None
    nrm = torch.empty((nrows,), dtype=torch.float32, device=dev) if want_norm else None
    d = gx_gather_desc()
    d.nlevels = len(feats_nhwc)
    for i, f in enumerate(feats_nhwc):
        _f32(f, "feature")
        d.feat[i] = f.data_ptr()
        d.h[i], d.w[i], d.c[i] = f.shape[1], f.shape[2], f.shape[3]
    d.out_h, d.out_w, d.hlen = out_h, out_w, hlen
    d.row_img, d.row_src, d.nrows = _ptr(row_img), _ptr(row_src), nrows
    d.a_hi, d.a_lo, d.a_f32, d.ld = _ptr(a_hi), _ptr(a_lo), _ptr(a_f), ld
    d.row_norm = _ptr(nrm)
    with timed("gather_rows", float(nrows) * hlen * (4 + (4 if want_lo else 2))):
        _check(lib.gx_gather_rows(C.byref(d), _stream()), "gx_gather_rows")
    _count()
    if want_norm:
        return a_hi, a_lo, a_f, nrm
    return a_hi, a_lo, a_f


def l2norm_split(z, want_lo=True, row_idx=None, want_f16=False):
    """zn = normalise(z[row_idx]) (row_idx int32 [n], -1 = zero row) or normalise(z).
    Returns (hi, lo, inv_norm) or, with want_f16, (hi, lo, inv_norm, fp16 plane)."""
    lib = load()
    _f32(z, "z")
    c = z.shape[1]
    n = z.shape[0] if row_idx is None else row_idx.numel()
    hi = torch.empty((n, c), dtype=torch.bfloat16, device=z.device)
    lo = torch.empty_like(hi) if want_lo else None
    inv = torch.empty((n,), dtype=torch.float32, device=z.device)
    f16 = torch.empty((n, c), dtype=torch.float16, device=z.device) if want_f16 else None
    with timed("l2norm_split", float(n) * c * (4 + (4 if want_lo else 2) + (2 if want_f16 else 0))):
        _check(lib.gx_l2norm_split(_ptr(z), _ptr(row_idx), _ptr(hi), _ptr(lo), _ptr(f16), _ptr(inv), n, c,
                                   _stream()), "gx_l2norm_split")
    _count()
    if want_f16:
        return hi, lo, inv, f16
    return hi, lo, inv


def l2norm_bwd_split(dzn, zn_hi, zn_lo, inv_norm, want_lo=True, want_planes=True, out_f32=None, out_hi=None):
    """dz planes and / or fp32 rows written into `out_f32` [n,c]; `out_hi`: a caller-provided bf16 [n,c] (slice) for
    the hi plane"""
    lib = load()
    _f32(dzn, "dzn")
    n, c = dzn.shape
    if out_hi is not None:
        if out_hi.dtype != torch.bfloat16 or tuple(out_hi.shape) != (n, c) or not out_hi.is_contiguous():
            raise GxError("l2norm_bwd_split: out_hi must be a contiguous bf16 [n, c] tensor")
        hi, want_planes = out_hi, True
    else:
        hi = torch.empty((n, c), dtype=torch.bfloat16, device=dzn.device) if want_planes else None
    lo = torch.empty_like(hi) if (want_lo and want_planes) else None
    if out_f32 is not None:
        _f32(out_f32, "out_f32")
    with timed("l2norm_bwd_split", float(n) * c * (4 + (4 if zn_lo is not None else 2) + (4 if want_lo else 2))):
        _check(lib.gx_l2norm_bwd_split(_ptr(dzn), _ptr(zn_hi), _ptr(zn_lo), _ptr(inv_norm), _ptr(hi), _ptr(lo),
                                       _ptr(out_f32), n, c, _stream()), "gx_l2norm_bwd_split")
    _count()
    return hi, lo


def segment_sum_rows(rows, order, seg_off, nseg, want_lo=False, want_planes=True, want_f32=False):
    """per-segment sums of `rows[order[...]]` as bf16 planes [nseg, c] and / or fp32"""
    lib = load()
    bf16_rows = rows.dtype == torch.bfloat16
    if not bf16_rows:
        _f32(rows, "rows")
    elif not rows.is_contiguous():
        raise GxError("segment_sum_rows: rows must be contiguous")
    c = rows.shape[1]
    hi = torch.empty((nseg, c), dtype=torch.bfloat16, device=rows.device) if want_planes else None
    lo = torch.empty_like(hi) if (want_lo and want_planes) else None
    f = torch.empty((nseg, c), dtype=torch.float32, device=rows.device) if want_f32 else None
    with timed("segment_sum_rows", float(rows.shape[0]) * c * (2 if bf16_rows else 4) +
               float(nseg) * c * (4 if want_lo or want_f32 else 2)):
        _check(lib.gx_segment_sum_rows(_ptr(rows), int(bf16_rows), _ptr(order), _ptr(seg_off), _ptr(hi), _ptr(lo), _ptr(f),
                                       nseg, c, _stream()), "gx_segment_sum_rows")
    _count()
    return hi, lo, f


def upsample_sum(parts, batch, out_h, out_w, out=None, planes=None, bilinear=False, labels=None):
    """parts: list of fp32 [batch, h_l, w_l, c] tensors -> fp32 [batch*out_h*out_w, c]; labels (int64
    [batch*out_h*out_w], optional): filled with the first arg-max over c of every pixel by the same kernel"""
    lib = load()
    n = len(parts)
    c = parts[0].shape[3]
    ptrs = (C.c_void_p * n)(*[p.data_ptr() for p in parts])
    hs = (C.c_int * n)(*[p.shape[1] for p in parts])
    ws = (C.c_int * n)(*[p.shape[2] for p in parts])
    for p in parts:
        _f32(p, "part")
    if out is None:
        out = torch.empty((batch * out_h * out_w, c), dtype=torch.float32, device=parts[0].device)
    assert out.is_contiguous() and out.numel() == batch * out_h * out_w * c
    if labels is not None and (labels.dtype != torch.int64 or labels.numel() != batch * out_h * out_w or
                               not labels.is_contiguous()):
        raise GxError("upsample_sum: labels must be a contiguous int64 tensor of batch*out_h*out_w elements")
    nbytes = 4.0 * c * (sum(p.shape[0] * p.shape[1] * p.shape[2] for p in parts) + batch * out_h * out_w)
    with timed("upsample_sum", nbytes):
        hi, lo = planes if planes is not None else (None, None)
        _check(lib.gx_upsample_sum(n, ptrs, hs, ws, batch, out_h, out_w, c, _ptr(out), _ptr(hi), _ptr(lo),
                                   _ptr(labels), int(bool(bilinear)), _stream()), "gx_upsample_sum")
    _count()
    return out


def tap_sum(g, batch, h, w, cout, dilation, bias, act, want_out=True, want_planes=False, want_lo=True):
    """g fp32 [batch*h*w, 9*cout] -> (out fp32 [batch,h,w,cout] or None, hi, lo [batch,h,w,pad64(cout)] or None)"""
    lib = load()
    _f32(g, "g"), _f32(bias, "bias")
    dev = g.device
    out = torch.empty((batch, h, w, cout), dtype=torch.float32, device=dev) if want_out else None
    hi = lo = None
    next_ld = pad64(cout)
    if want_planes:
        hi, lo = _planes((batch, h, w, next_ld), dev, next_ld != cout, want_lo)
    with timed("tap_sum", 4.0 * batch * h * w * cout * 10):
        _check(lib.gx_tap_sum(_ptr(g), batch, h, w, cout, int(dilation), _ptr(bias), int(act), _ptr(out), _ptr(hi),
                              _ptr(lo), next_ld, _stream()), "gx_tap_sum")
    _count()
    return out, hi, lo


def tap_spread(dout_nhwc, dilation, want_lo=True):
    """dout fp32 [b,h,w,cout] -> bf16 planes [b*h*w, 9*cout] (adjoint of tap_sum)"""
    lib = load()
    _f32(dout_nhwc, "dout")
    b, h, w, cout = dout_nhwc.shape
    hi = torch.empty((b * h * w, 9 * cout), dtype=torch.bfloat16, device=dout_nhwc.device)
    lo = torch.empty_like(hi) if want_lo else None
    with timed("tap_spread", float(b) * h * w * cout * (4 + 9 * (4 if want_lo else 2))):
        _check(lib.gx_tap_spread(_ptr(dout_nhwc), b, h, w, cout, int(dilation), _ptr(hi), _ptr(lo), _stream()),
               "gx_tap_spread")
    _count()
    return hi, lo


def pool_bilinear_adjoint(x_nhwc, out_h, out_w):
    """adjoint of F.interpolate(mode='bilinear', align_corners=False) from [b,out_h,out_w,c] to x's resolution:
    fp32 [b,H,W,c] -> fp32 [b,out_h,out_w,c] (x then y, separable)"""
    lib = load()
    _f32(x_nhwc, "x")
    b, h, w, c = x_nhwc.shape
    dev = x_nhwc.device
    t = x_nhwc
    if out_w != w:
        t2 = torch.empty((b, h, out_w, c), dtype=torch.float32, device=dev)
        with timed("pool_bilinear", 4.0 * b * h * c * (w + out_w)):
            _check(lib.gx_pool1d_bilinear(_ptr(t), b * h, w, out_w, c, _ptr(t2), _stream()), "gx_pool1d_bilinear")
        _count()
        t = t2
    if out_h != h:
        t2 = torch.empty((b, out_h, out_w, c), dtype=torch.float32, device=dev)
        with timed("pool_bilinear", 4.0 * b * out_w * c * (h + out_h)):
            _check(lib.gx_pool1d_bilinear(_ptr(t), b, h, out_h, out_w * c, _ptr(t2), _stream()),
                   "gx_pool1d_bilinear")
        _count()
        t = t2
    return t


def pool_sum(x_nhwc, out_h, out_w, want_f32=True, want_planes=True, want_lo=False):
    """block sums of fp32 [b,H,W,c] onto [b,out_h,out_w,c]: (fp32, hi, lo)"""
    lib = load()
    _f32(x_nhwc, "x")
    b, h, w, c = x_nhwc.shape
    dev = x_nhwc.device
    f = torch.empty((b, out_h, out_w, c), dtype=torch.float32, device=dev) if want_f32 else None
    hi = torch.empty((b * out_h * out_w, c), dtype=torch.bfloat16, device=dev) if want_planes else None
    lo = torch.empty_like(hi) if (want_lo and want_planes) else None
    with timed("pool_sum", 4.0 * b * c * (h * w + out_h * out_w)):
        _check(lib.gx_pool_sum(_ptr(x_nhwc), b, h, w, out_h, out_w, c, _ptr(f), _ptr(hi), _ptr(lo), _stream()),
               "gx_pool_sum")
    _count()
    return f, hi, lo


def normalize_rows_(w):
    lib = load()
    _f32(w, "w")
    _check(lib.gx_normalize_rows(_ptr(w), w.shape[0], w.shape[1], _stream()), "gx_normalize_rows")
    _count()
    return w


class SinkhornWorkspace:
    def __init__(self, k, device):
        lib = load()
        self.k = k
        self.max_parts = lib.gx_sinkhorn_max_parts()
        self.partials = torch.empty((self.max_parts, k), dtype=torch.float32, device=device)
        self.u = torch.empty((k,), dtype=torch.float32, device=device)
        self._cache = {}

    def cache16(self, chain, n):
        """(e16 [n, lde] fp16, la1 [k]) of Sinkhorn chain `chain`: the 16-bit plane of gx_sinkhorn_pass_cached, kept
        between calls (1.6 GB per chain at n = 160000, k = 5000)"""
        lde = (self.k + 7) // 8 * 8
        ent = self._cache.get(chain)
        if ent is None or ent[0].shape[0] < n:
            dev = self.partials.device
            ent = (torch.empty((n, lde), dtype=torch.float16, device=dev),
                   torch.empty((lde,), dtype=torch.float32, device=dev))
            self._cache[chain] = ent
        return ent[0][:n], ent[1]


class LLExchange:
    """Exchange buffers of the distributed Sinkhorn on one box (gx_ll_desc, include/ganecdotes_b200.h): one
    cudaMalloc'ed buffer per rank, mapped into every rank with CUDA IPC; `channels` independent chains (the s and t
    views), each with two alternating blocks of [world][k] tagged words.  torch.distributed only carries the
    64-byte IPC handles at construction; the data path is plain NVLink stores issued by the kernels."""

    def __init__(self, pg, rank, world, k, device, channels=2):
        import torch.distributed as dist
        lib = load()
        if world > GX_MAX_PEERS:
            raise GxError(f"LLExchange: at most {GX_MAX_PEERS} ranks")
        if k % 4:
            raise GxError("LLExchange: k must be a multiple of 4")
        self.rank, self.world, self.k, self.channels = rank, world, k, channels
        self.block = world * k                                    # words per block
        nbytes = channels * 2 * self.block * 8
        self.own, self.ptrs = None, []
        # every step below that can fail locally is followed by a collective that tells the peers, so that no rank is
        # left waiting in a collective the failed rank never enters
        handle, local_err = None, None
        try:
            own = C.c_void_p()
            _check(lib.gx_peer_alloc(nbytes, C.byref(own)), "gx_peer_alloc")
            self.own = own.value
            h = (C.c_char * 64)()
            _check(lib.gx_peer_export(self.own, h), "gx_peer_export")
            handle = bytes(h)
        except Exception as e:          # noqa: BLE001
            local_err = e
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=pg)
        if any(h is None for h in handles):
            self.close()
            raise GxError(f"LLExchange: peer buffer allocation / export failed on a rank ({local_err!r})")
        open_err = None
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own)
                continue
            p = C.c_void_p()
            buf = (C.c_char * 64).from_buffer_copy(handles[r])
            rc = lib.gx_peer_open(buf, C.byref(p))
            if rc != 0:
                open_err = GxError(f"gx_peer_open failed for rank {r} (rc {rc}, cuda error {lib.gx_last_cuda_error()})")
                self.ptrs.append(None)
            else:
                self.ptrs.append(p.value)
        flags = [None] * world
        dist.all_gather_object(flags, open_err is None, group=pg)     # also: every buffer is zeroed and mapped from here on
        if not all(flags):
            self.close()
            raise open_err or GxError("LLExchange: a peer could not map the exchange buffers")
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.seq = [0] * channels

    @classmethod
    def simulated(cls, world, k, device, channels=2):
        """`world` endpoints inside ONE process (all buffers on `device`, no IPC): exercises the slot / parity /
        sequence protocol of the kernels on a single GPU.  The caller must issue every endpoint's send before any
        endpoint's receive on the stream (a receive spins until all sends of its exchange have landed)."""
        lib = load()
        ptrs = []
        for _ in range(world):
            p = C.c_void_p()
            _check(lib.gx_peer_alloc(channels * 2 * world * k * 8, C.byref(p)), "gx_peer_alloc")
            ptrs.append(p.value)
        ends = []
        for r in range(world):
            e = cls.__new__(cls)
            e.rank, e.world, e.k, e.channels, e.block = r, world, k, channels, world * k
            e.ptrs, e.own = list(ptrs), None         # buffers are owned by endpoint 0's `_sim_owned`
            e.err = torch.zeros(1, dtype=torch.int32, device=device)
            e.seq = [0] * channels
            ends.append(e)
        ends[0]._sim_owned = ptrs
        return ends

    def _desc(self, channel, seq):
        d = gx_ll_desc()
        for r in range(self.world):
            d.peers[r] = self.ptrs[r]
        d.world, d.rank = self.world, self.rank
        d.block_words = (channel * 2 + (seq & 1)) * self.block
        d.seq = seq
        d.err = self.err.data_ptr()
        return d

    def next_send(self, channel):
        """descriptor of the next exchange of `channel` (advances its sequence number)"""
        self.seq[channel] += 1
        return self._desc(channel, self.seq[channel])

    def last(self, channel):
        """descriptor of the most recent exchange of `channel` (what a consumer receives)"""
        return self._desc(channel, self.seq[channel])

    def check(self):
        """raises if a consumer gave up waiting for a peer (synchronises)"""
        if int(self.err.item()) != 0:
            raise GxError("LLExchange: a rank did not receive its peers' marginals within the time-out")

    def close(self):
        lib = load()
        if getattr(self, "_sim_owned", None):
            for p in self._sim_owned:
                lib.gx_peer_free(p)
            self._sim_owned, self.ptrs = None, []
            return
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                lib.gx_peer_close(p)
        if self.own:
            lib.gx_peer_free(self.own)
        self.ptrs, self.own = [], None


def sinkhorn_pass_parts(s, inv_eps, first, u_in, r, c, n_total, ws: SinkhornWorkspace, u_ll=None, reverse=False):
    """One streaming pass over S [n,k]: per-CTA partial column sums in ws.partials; returns their count."""
    lib = load()
    n, k = s.shape
    nparts = C.c_int(0)
    with timed("sinkhorn_pass", 4.0 * n * k):
        _check(lib.gx_sinkhorn_pass(_ptr(s), n, k, s.stride(0), float(inv_eps), int(first), _ptr(u_in),
                                    C.byref(u_ll) if u_ll is not None else None, _ptr(r), _ptr(c), int(n_total),
                                    int(bool(reverse)), _ptr(ws.partials), C.byref(nparts), _stream()),
               "gx_sinkhorn_pass")
    _count()
    return nparts.value


def sinkhorn_pass_cached_parts(s, inv_eps, u_in, r, c, n_total, ws: SinkhornWorkspace, cache, write_cache, u_ll=None,
                               reverse=False):
    """A (non-first) pass through the 16-bit cache `cache` = ws.cache16(chain, n): write_cache=True streams S and fills
    the cache, write_cache=False streams the cache instead of S (gx_sinkhorn_pass_cached)."""
    lib = load()
    n, k = s.shape
    e16, la1 = cache
    nparts = C.c_int(0)
    with timed("sinkhorn_pass", (4.0 * n * k + 2.0 * n * k) if write_cache else 2.0 * n * k):
        _check(lib.gx_sinkhorn_pass_cached(_ptr(s), n, k, s.stride(0), float(inv_eps), _ptr(u_in),
                                           C.byref(u_ll) if u_ll is not None else None, _ptr(r), _ptr(c), int(n_total),
                                           int(bool(reverse)), _ptr(ws.partials), C.byref(nparts), _ptr(e16),
                                           e16.stride(0), _ptr(la1), int(bool(write_cache)), _stream()),
               "gx_sinkhorn_pass_cached")
    _count()
    return nparts.value


def sinkhorn_reduce(parts, nparts, k, out):
    _check(load().gx_sinkhorn_reduce(_ptr(parts), int(nparts), int(k), _ptr(out), _stream()), "gx_sinkhorn_reduce")
    _count()
    return out


def sinkhorn_reduce_send(parts, nparts, k, ll_desc, u_local=None):
    """column sums of the partials, pushed into this rank's slot of the exchange block on every rank"""
    _check(load().gx_sinkhorn_reduce_send(_ptr(parts), int(nparts), int(k), C.byref(ll_desc), _ptr(u_local), _stream()),
           "gx_sinkhorn_reduce_send")
    _count()


def ll_recv_sum(ll_desc, k, device):
    u = torch.empty((k,), dtype=torch.float32, device=device)
    _check(load().gx_ll_recv_sum(C.byref(ll_desc), int(k), _ptr(u), _stream()), "gx_ll_recv_sum")
    _count()
    return u


def sinkhorn_pass(s, inv_eps, first, u_in, r, c, n_total, ws: SinkhornWorkspace, reverse=False):
    """One streaming pass over S [n,k]; afterwards ws.u holds the LOCAL column sums."""
    nparts = sinkhorn_pass_parts(s, inv_eps, first, u_in, r, c, n_total, ws, reverse=reverse)
    return sinkhorn_reduce(ws.partials, nparts, s.shape[1], ws.u)


def sinkhorn_log_a(u, r, u_ll=None, k=None, device=None):
    lib = load()
    k = u.numel() if u is not None else k
    la = torch.empty((k,), dtype=torch.float32, device=u.device if u is not None else device)
    _check(lib.gx_sinkhorn_log_a(_ptr(u), C.byref(u_ll) if u_ll is not None else None, _ptr(r), k, _ptr(la),
                                 _stream()), "gx_sinkhorn_log_a")
    _count()
    return la


def sinkhorn_q(s, inv_eps, log_a):
    lib = load()
    n, k = s.shape
    q = torch.empty((n, k), dtype=torch.float32, device=s.device)
    _check(lib.gx_sinkhorn_q(_ptr(s), n, k, s.stride(0), float(inv_eps), _ptr(log_a), _ptr(q), _stream()),
           "gx_sinkhorn_q")
    _count()
    return q


def swav_loss(s_s, s_t, inv_eps, inv_temp, la_s, la_t, grad_scale, want_lo=False, want_f32=False, want_db=True,
              loss_parts=None, db_accum=None):
    """Returns (loss_parts [gx_loss_max_parts()] per-CTA sums (not divided by N), dS_s planes, dS_t planes, db [k],
    f32 grads).  loss_parts: a zeroed caller buffer to fill (else allocated); db_accum: the bias gradient is
    accumulated into it (db_accum += column sums of dS_s + dS_t) instead of being returned."""
    lib = load()
    n, k = s_s.shape
    dev = s_s.device
    maxp = lib.gx_loss_max_parts()
    if loss_parts is None:
        loss_parts = torch.zeros((maxp,), dtype=torch.float32, device=dev)
    db_parts = torch.empty((maxp, k), dtype=torch.float32, device=dev) if want_db else None
    ds_s_hi = torch.empty((n, k), dtype=torch.bfloat16, device=dev)
    ds_t_hi = torch.empty((n, k), dtype=torch.bfloat16, device=dev)
    ds_s_lo = torch.empty_like(ds_s_hi) if want_lo else None
    ds_t_lo = torch.empty_like(ds_t_hi) if want_lo else None
    fs = torch.empty((n, k), dtype=torch.float32, device=dev) if want_f32 else None
    ft = torch.empty((n, k), dtype=torch.float32, device=dev) if want_f32 else None
    nparts = C.c_int(0)
    with timed("swav_loss_fwd_bwd", 2.0 * n * k * (4 + (4 if want_lo else 2))):
        _check(lib.gx_swav_loss(_ptr(s_s), _ptr(s_t), n, k, s_s.stride(0), float(inv_eps), float(inv_temp),
                                _ptr(la_s), _ptr(la_t), float(grad_scale), _ptr(loss_parts), _ptr(db_parts),
                                C.byref(nparts), _ptr(ds_s_hi), _ptr(ds_s_lo), _ptr(ds_t_hi), _ptr(ds_t_lo), k,
                                _ptr(fs), _ptr(ft), _stream()), "gx_swav_loss")
    _count()
    db = None
    if want_db and db_accum is not None:
        colsum(db_parts, nparts.value, k, db_accum, 1.0, True)
    elif want_db:
        db = torch.empty((k,), dtype=torch.float32, device=dev)
        _check(lib.gx_sinkhorn_reduce(_ptr(db_parts), nparts.value, k, _ptr(db), _stream()), "gx_sinkhorn_reduce")
        _count()
    return loss_parts, (ds_s_hi, ds_s_lo), (ds_t_hi, ds_t_lo), db, (fs, ft)


def im2col(x, kh, kw, stride, padding, dilation, ho, wo, ld):
    """x [b,c,h,w] fp32 -> cols [b*ho*wo, ld] fp32 (gx_im2col)"""
    _f32(x, "x")
    b, c, h, w = x.shape
    cols = torch.empty((b * ho * wo, ld), dtype=torch.float32, device=x.device)
    _check(load().gx_im2col(_ptr(x), b, c, h, w, kh, kw, stride[0], stride[1], padding[0], padding[1], dilation[0],
                            dilation[1], ho, wo, ld, _ptr(cols), _stream()), "gx_im2col")
    _count()
    return cols


def col2im(cols, shape, kh, kw, stride, padding, dilation, ho, wo):
    """cols [b*ho*wo, ld] fp32 -> x [b,c,h,w] fp32: the adjoint of im2col (gx_col2im)"""
    _f32(cols, "cols")
    b, c, h, w = shape
    x = torch.empty((b, c, h, w), dtype=torch.float32, device=cols.device)
    _check(load().gx_col2im(_ptr(cols), b, c, h, w, kh, kw, stride[0], stride[1], padding[0], padding[1], dilation[0],
                            dilation[1], ho, wo, cols.stride(0), _ptr(x), _stream()), "gx_col2im")
    _count()
    return x


def recip_clamp(x, eps):
    """1 / max(x, eps)"""
    _f32(x, "x")
    out = torch.empty_like(x)
    _check(load().gx_recip(_ptr(x), x.numel(), float(eps), 0, _ptr(out), _stream()), "gx_recip")
    _count()
    return out


def rsqrt_eps(x, eps):
    """rsqrt(x + eps)"""
    _f32(x, "x")
    out = torch.empty_like(x)
    _check(load().gx_recip(_ptr(x), x.numel(), float(eps), 1, _ptr(out), _stream()), "gx_recip")
    _count()
    return out


def bn_stats(hraw, rscale, eps=1e-5, run_mean=None, run_var=None, momentum=0.1):
    """(mean[c], invstd[c]) of h = hraw * rscale[row] over the rows (training-mode BatchNorm1d); updates the
    running statistics in place when given"""
    _f32(hraw, "hraw"), _f32(rscale, "rscale")
    n, c = hraw.shape
    mean = torch.empty((c,), dtype=torch.float32, device=hraw.device)
    invstd = torch.empty_like(mean)
    _check(load().gx_bn_stats(_ptr(hraw), hraw.stride(0), _ptr(rscale), n, c, float(eps), _ptr(mean), _ptr(invstd),
                              _ptr(run_mean), _ptr(run_var), float(momentum), _stream()), "gx_bn_stats")
    _count()
    return mean, invstd


def bn_act_apply(hraw, rscale, mean, invstd, gamma, beta, slope=0.01, want_f32=True, want_planes=True, want_lo=True):
    """a = lrelu((hraw * rscale - mean) * invstd * gamma + beta) -> (a fp32 or None, hi, lo)"""
    _f32(hraw, "hraw"), _f32(rscale, "rscale")
    n, c = hraw.shape
    dev = hraw.device
    out = torch.empty((n, c), dtype=torch.float32, device=dev) if want_f32 else None
    hi = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if want_planes else None
    lo = torch.empty_like(hi) if (want_planes and want_lo) else None
    with timed("bn_act_apply", float(n) * c * (4 + (4 if want_f32 else 0) + (4 if lo is not None else 2 if hi is not None else 0))):
        _check(load().gx_bn_act_apply(_ptr(hraw), hraw.stride(0), _ptr(rscale), n, c, _ptr(mean), _ptr(invstd), _ptr(gamma),
                                      _ptr(beta), float(slope), _ptr(out), _ptr(hi), _ptr(lo), _stream()), "gx_bn_act_apply")
    _count()
    return out, hi, lo


def bn_act_bwd(da, hraw, rscale, mean, invstd, gamma, beta, slope=0.01):
    """(dhs [n,c] = dL/dhraw, dgamma [c], dbeta [c])"""
    _f32(da, "da"), _f32(hraw, "hraw")
    n, c = hraw.shape
    dhs = torch.empty((n, c), dtype=torch.float32, device=hraw.device)
    dg = torch.empty((c,), dtype=torch.float32, device=hraw.device)
    db = torch.empty_like(dg)
    _check(load().gx_bn_act_bwd(_ptr(da), _ptr(hraw), hraw.stride(0), _ptr(rscale), n, c, _ptr(mean), _ptr(invstd),
                                _ptr(gamma), _ptr(beta), float(slope), _ptr(dhs), _ptr(dg), _ptr(db), _stream()),
           "gx_bn_act_bwd")
    _count()
    return dhs, dg, db


def simclr_loss(z, temperature):
    """(loss [1], dz [n2, c]) of the reference's contrastive loss on the projection output z [n2, c]"""
    _f32(z, "z")
    n2, c = z.shape
    loss = torch.empty((1,), dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z)
    _check(load().gx_simclr_loss(_ptr(z), n2, c, 1.0 / float(temperature), _ptr(loss), _ptr(dz), _stream()),
           "gx_simclr_loss")
    _count()
    return loss, dz


def larc_scratch(device):
    return torch.empty(load().gx_larc_scratch_floats(), dtype=torch.float32, device=device)


def larc_sgd_(p, g, buf, lr, momentum, trust, weight_decay, eps, first_step, norms):
    lib = load()
    if norms.numel() < lib.gx_larc_scratch_floats():
        raise GxError("larc_sgd_: norms scratch must hold gx_larc_scratch_floats() floats (L.larc_scratch)")
    _f32(p, "p"), _f32(g, "g"), _f32(buf, "buf")
    _check(lib.gx_larc_sgd(_ptr(p), _ptr(g), _ptr(buf), p.numel(), float(lr), float(momentum), float(trust),
                           float(weight_decay), float(eps), int(first_step), _ptr(norms), _stream()), "gx_larc_sgd")
    _count(2)


def argmax_rows(x, out=None):
    lib = load()
    n, c = x.shape
    labels = out if out is not None else torch.empty((n,), dtype=torch.int64, device=x.device)
    if labels.dtype != torch.int64 or labels.numel() != n or not labels.is_contiguous():
        raise GxError("argmax_rows: out must be a contiguous int64 tensor of n elements")
    _check(lib.gx_argmax_rows(_ptr(x), n, c, x.stride(0), _ptr(labels), _stream()), "gx_argmax_rows")
    _count()
    return labels


def argmin_affine(s, bias, scale):
    """labels int32 [n] = first argmin_k (bias[k] + scale * s[n,k])"""
    lib = load()
    _f32(s, "s"), _f32(bias, "bias")
    n, k = s.shape
    labels = torch.empty((n,), dtype=torch.int32, device=s.device)
    with timed("kmeans_argmin", 4.0 * n * k):
        _check(lib.gx_argmin_affine(_ptr(s), n, k, s.stride(0), _ptr(bias), float(scale), _ptr(labels), _stream()),
               "gx_argmin_affine")
    _count()
    return labels


_CENTER_PLANES = {}


def _center_planes(centers):
    """split-bf16 planes and squared norms of a centre matrix, cached per (storage, version): the centres of a fitted
    model are re-used for every image"""
    key = (centers.data_ptr(), centers._version, tuple(centers.shape), tuple(centers.stride()), str(centers.device))
    hit = _CENTER_PLANES.get(key)
    if hit is not None and hit[-1] is not centers and hit[-1].untyped_storage().data_ptr() != \
            centers.untyped_storage().data_ptr():
        hit = None
    if hit is None:
        if len(_CENTER_PLANES) > 64:
            _CENTER_PLANES.clear()
        c_hi, c_lo = split_planes(centers.contiguous())
        cn = (centers.double() ** 2).sum(1).float().contiguous()        # ||c_k||^2: K numbers, once per model
        frags = cn_pad = None
        k, c = centers.shape
        nbytes = load().gx_kmeans_frag_bytes(k, c)
        if nbytes > 0:      # fragments of the fused kernel (k <= 64): centres in the order its lanes read them
            frags = torch.empty((nbytes // 16, 4), dtype=torch.int32, device=centers.device)
            _check(load().gx_kmeans_center_frags(_ptr(centers.contiguous()), k, c, _ptr(frags), _stream()),
                   "gx_kmeans_center_frags")
            _count()
            cn_pad = torch.full((64,), float("inf"), dtype=torch.float32, device=centers.device)
            cn_pad[:k] = cn
        # the entry keeps `centers` alive: its address cannot be handed to another tensor while the entry exists
        hit = _CENTER_PLANES[key] = (c_hi, c_lo, cn, frags, cn_pad, centers)
    return hit


def kmeans_assign(x, centers, x2=None, want_dist=False, tensor=None):
    """x [n,c1] (+ x2 [n,c2]) fp32 contiguous, centers [k,c1+c2] -> int32 labels [n]
    (with want_dist: (labels, squared distance to the assigned centre [n])).

    Two routes.  Direct (SIMT, sum of (x - c)^2 in fp32): exact distances, used for the k-means fit (want_dist) and
    for small inputs.  Tensor cores (`tensor`, default for labels-only calls on >= 4096 rows): scores X C^T from
    split-bf16 operands (fp32-grade products), then argmin_k(||c_k||^2 - 2 x.c_k) - the GEMM form scikit-learn's
    predict uses; the features are read once instead of once per centre.  For k <= 64 the whole route is ONE fused
    kernel (`gx_kmeans_assign_mma`: rows split in registers, mma.sync against centre fragments in shared memory;
    channels % 16 == 0); larger k (or tensor="gemm") goes through split_planes -> gx_gemm (3 passes) ->
    gx_argmin_affine.  The two routes can differ only where the
    two nearest centres are closer than ~1e-4 of the squared distance (accumulation rounding of the x.c term)."""
    lib = load()
    _f32(centers, "centers"), _f32(x, "x"), _f32(x2, "x2")
    n, c1 = x.shape
    c2 = 0 if x2 is None else x2.shape[1]
    c = c1 + c2
    if centers.shape[1] != c:
        raise GxError("kmeans_assign: centers must have c1+c2 columns")
    if tensor is None:
        tensor = (not want_dist) and n >= 4096 and c1 % 8 == 0 and c2 % 8 == 0
    if tensor:
        if want_dist:
            raise GxError("kmeans_assign: the tensor-core route returns labels only")
        k = centers.shape[0]
        if tensor != "gemm" and c1 % 16 == 0 and c2 % 16 == 0:
            frags, cn_pad = _center_planes(centers)[3:5]
            if frags is not None:       # fused route: rows read once, no planes / scores in HBM
                labels = torch.empty((n,), dtype=torch.int32, device=x.device)
                with timed("kmeans_assign_fused", 4.0 * n * c):
                    _check(lib.gx_kmeans_assign_mma(_ptr(x), c1, _ptr(x2), c2, n, _ptr(frags), _ptr(cn_pad), k,
                                                    _ptr(labels), _stream()), "gx_kmeans_assign_mma")
                _count()
                return labels
        a_hi = torch.empty((n, c), dtype=torch.bfloat16, device=x.device)
        a_lo = torch.empty_like(a_hi)
        split_planes(x, out=(a_hi[:, :c1], a_lo[:, :c1]))
        if x2 is not None:
            split_planes(x2, out=(a_hi[:, c1:], a_lo[:, c1:]))
        c_hi, c_lo, cn = _center_planes(centers)[:3]
        s = gemm(a_hi, a_lo, c_hi, c_lo, n, k, c, 3, tag="gemm_kmeans_scores", block_n=64 if k <= 64 else 0)
        return argmin_affine(s, cn, -2.0)
    labels = torch.empty((n,), dtype=torch.int32, device=x.device)
    dist = torch.empty((n,), dtype=torch.float32, device=x.device) if want_dist else None
    with timed("kmeans_assign", 4.0 * n * c):
        _check(lib.gx_kmeans_assign(_ptr(x), c1, _ptr(x2), c2, n, _ptr(centers), centers.shape[0], _ptr(labels),
                                    _ptr(dist), _stream()), "gx_kmeans_assign")
    _count()
    return (labels, dist) if want_dist else labels


def onehot_nearest(labels_bhw, k, out_h, out_w, out=None, on=1.0, off=0.0):
    """one-hot maps [b, k, out_h, out_w] of int32 labels [b, h, w], nearest resize; `out`: a channel slice
    [b, k, out_h, out_w] of a wider contiguous [b, K_total, out_h, out_w] tensor to write into"""
    lib = load()
    b, h, w = labels_bhw.shape
    if out is None:
        out = torch.empty((b, k, out_h, out_w), dtype=torch.float32, device=labels_bhw.device)
    elif (out.dtype != torch.float32 or tuple(out.shape) != (b, k, out_h, out_w) or out.stride(3) != 1 or
          out.stride(2) != out_w or out.stride(1) != out_h * out_w):
        raise GxError("onehot_nearest: out must be a channel slice of a contiguous [b, K, out_h, out_w] float tensor")
    with timed("onehot_nearest", 4.0 * b * k * out_h * out_w):
        _check(lib.gx_onehot_nearest(_ptr(labels_bhw), b, h, w, k, out_h, out_w, _ptr(out),
                                     out.stride(0) if b > 1 else 0, float(on), float(off), _stream()), "gx_onehot_nearest")
    _count()
    return out
