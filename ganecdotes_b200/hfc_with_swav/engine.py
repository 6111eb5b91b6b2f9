"""Fused SwAV engine: the batched / sharded hot path behind `SwAVClustering`.

One optimiser step (ref hfc_with_swav/swav_clustering.py:320-460) for a batch of B
latents on this rank (joint-batch = SwAV "distributed Sinkhorn" semantics; B = 1 on one
rank is exactly the reference):

  w = style(z)                                   mapping MLP kernels
  per view:  W+ with two perturbed rows, double truncation        (ref :593-640, aug:42-53,75-79)
             synthesis (tcgen05 implicit-GEMM convs, fused FIR)   -> 13 NHWC feature maps
  per patch: gather(rotate+flip+sample+upsample+concat)           -> A [B*N, hlen]  (bf16 planes)
             Z = A Wp^T ; Zn = Z/|Z| ; S = Zn Wk^T + b            (tcgen05 GEMMs)
             Sinkhorn: niters streaming passes over S, only the K-vector of column
                       marginals is exchanged (all-reduce over ranks)
             fused swapped-prediction loss fwd + dS
             dZn = dS Wk ; gWk += dS^T Zn ; dZ = normalise'(dZn) ; gWp += dZ^T A
  all-reduce gradients ; LARC + SGD(momentum)

Everything on the device is a hand-written kernel from `ganecdotes_b200._lib`; the only
host arithmetic is the data-independent index bookkeeping (rotation/flip index map,
random permutations), a few KB per latent.
"""
import math
import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import _lib as L


# ----------------------------------------------------------------------------------------
# host-side index bookkeeping
# ----------------------------------------------------------------------------------------

def rotate_flip_index_map(h: int, w: int, angle: float, flip: bool) -> torch.Tensor:
    """int64 [h*w] on the CPU: for every pixel of `flip(rotate(x, angle))` the flat index of
    the source pixel it copies, or -1 where the rotation fills with zeros.

    The map is produced by pushing an index image through torchvision's own rotate / hflip
    (the ops `fixed_transforms` applies to the feature tensor, ref swav_clustering.py:98-102,
    358-359), so the nearest-neighbour rounding is torchvision's, bit for bit."""
    import torchvision.transforms.functional as TF
    from torchvision.transforms import InterpolationMode
    idx = (torch.arange(h * w, dtype=torch.float32) + 1).view(1, 1, h, w)
    y = TF.rotate(idx, float(angle), InterpolationMode.NEAREST, False, None, [0.0])
    if flip:
        y = TF.hflip(y)
    return y.round().long().flatten() - 1


@dataclass
class ViewDraws:
    """Random draws of one view for every latent of the local batch."""
    layer_no: List[int]                 # np.random.choice(n_layers) per latent   (ref :610-612)
    pert_z: torch.Tensor                # [B, 2*n_layers, D] randn_like draws     (ref aug:47)
    angle: List[float]                  # RandomRotation angle per latent        (ref :358-359)
    flip: List[bool]                    # RandomHorizontalFlip decision per latent


@dataclass
class StepDraws:
    z: torch.Tensor                     # [B, D] latents (ref :323)
    view_s: ViewDraws
    view_t: ViewDraws
    perms: List[List[torch.Tensor]]     # perms[p][b]: randperm(H*W) (ref :388), shared by both views


_MAP_POOL = None


def _map_pool():
    global _MAP_POOL
    if _MAP_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _MAP_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="gx-index-map")
    return _MAP_POOL


def patch_pick_rows(h: int, w: int, pick: int, patch_size: int) -> torch.Tensor:
    """Pixel indices (row-major, int64) of the square crop [pick:pick+P, pick:pick+P] - what
    sampling_method == 'patch' feeds the projection (ref swav_clustering.py:150-158: the same offset on both
    axes, slices clipped at the border like Python's).  Used in place of a permutation: the P*P rows of a patch."""
    ys = torch.arange(pick, min(pick + patch_size, h))
    xs = torch.arange(pick, min(pick + patch_size, w))
    return (ys[:, None] * w + xs[None, :]).reshape(-1)


def build_row_indices(h, w, view: ViewDraws, perms, patch_size, device, count_fill=False):
    """[P, B*N] int32 (row_src), [B*N] int32 (row_img) for one view; with count_fill also the number
    of samples that fall on rotation fill (row_src == -1), counted on the host."""
    b = len(view.angle)
    if b >= 4:      # the torchvision ops release the GIL: a few host threads hide most of the map cost
        maps = list(_map_pool().map(lambda i: rotate_flip_index_map(h, w, view.angle[i], view.flip[i]), range(b)))
    else:
        maps = [rotate_flip_index_map(h, w, view.angle[i], view.flip[i]) for i in range(b)]
    n = patch_size if patch_size is not None else h * w
    rows = []
    for p in range(len(perms)):
        rows.append(torch.cat([maps[i][perms[p][i][:n]] for i in range(b)]))
    row_src = torch.stack(rows).to(torch.int32)
    row_img = torch.arange(b, dtype=torch.int32).repeat_interleave(n)
    if count_fill:
        return _upload(row_src, device), _upload(row_img, device), int((row_src < 0).sum())
    return _upload(row_src, device), _upload(row_img, device)


def _upload(t, device):
    """host -> device through pinned memory (truly asynchronous on the current stream)"""
    if torch.device(device).type != "cuda":
        return t.to(device)
    return t.contiguous().pin_memory().to(device, non_blocking=True)


# ----------------------------------------------------------------------------------------
# W+ construction for a perturbed view
# ----------------------------------------------------------------------------------------

@torch.no_grad()
def view_wplus(gen, w, mean_latent, truncation, view: ViewDraws, perturb_std):
    """W+ [B, n_latent, D] fed to the synthesis network for one view, including the
    reference's double truncation.  Only the two perturbed rows need
    the mapping network: the other ten `style(randn)` passes of the reference are
    multiplied by sigma = 0 (ref aug:42-53) and are skipped."""
    b, d = w.shape
    mean = mean_latent.reshape(-1).float().contiguous()
    wt = L.truncate(w.float().contiguous(), mean, truncation) if truncation < 1 else w
    wplus = wt.unsqueeze(1).repeat(1, gen.n_latent, 1).contiguous()
    rows = []
    for i in range(b):
        l = view.layer_no[i]
        rows += [view.pert_z[i, 2 * l], view.pert_z[i, 2 * l + 1]]
    noise_w = gen.style(torch.stack(rows).to(w.device).float().contiguous())       # [2B, D]
    for i in range(b):
        l = view.layer_no[i]
        sg = float(perturb_std[l])
        for j, r in enumerate((2 * l, 2 * l + 1)):
            wplus[i, r] = (1 - sg) * wplus[i, r] + sg * noise_w[2 * i + j]
    if truncation < 1:
        wplus = L.truncate(wplus, mean, truncation)                                   # second truncation
    return wplus


def regroup_order(nfeat: int) -> List[int]:
    """Feature order after the reference's 13 -> 7 regrouping (ref aug:80-90): it only
    concatenates neighbours, so the channel order of the final per-pixel vector is the
    plain layer order."""
    return list(range(nfeat))


# ----------------------------------------------------------------------------------------
# head state
# ----------------------------------------------------------------------------------------

def pick_split_k(tiles: int, kiters: int, sms: int = 148, min_iters: int = 8, max_split: int = 64) -> int:
    best, best_eff = 1, 0.0
    for s in range(1, max(1, min(max_split, kiters // min_iters)) + 1):
        work = tiles * s
        eff = work / (sms * math.ceil(work / sms))
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
    return best


class SwavHead:
    """Projection (Linear hlen->C, no bias) + prototype (Linear C->K with bias) layers,
    their bf16 operand planes, gradients and LARC/SGD state.  Operates IN PLACE on the
    parameters of the nn.Modules the caller saves (ref :504-505)."""

    def __init__(self, w_proj, w_proto, b_proto, lr, momentum, trust, passes_fwd=3, passes_bwd=1, proto_f16=False,
                 weight_decay=0.0, proj_slope=None):
        self.w_proj, self.w_proto, self.b_proto = w_proj, w_proto, b_proto
        # projn_nw == '1-layer' (ref :250-256): LeakyReLU(proj_slope) after the projection.  Point-wise, so the
        # project-every-pixel-once identity holds with the activation applied to Z; no shipped config uses it, and
        # it runs as two extra element-wise passes (`gx_fused_bias_act` forward / grad mode 1) rather than fused
        self.proj_slope = None if proj_slope is None else float(proj_slope)
        self.lr, self.momentum, self.trust, self.weight_decay = lr, momentum, trust, float(weight_decay)
        self.passes_fwd, self.passes_bwd = passes_fwd, passes_bwd
        # score GEMM operands.  Default: the 3-plane bf16 split, |dS| = 2e-7 rms -> codes Q within 5e-5 rms of
        # fp64 at the shipped eps = 0.005 (scores are multiplied by 200 inside the exponential).
        # proto_f16=True: single fp16 planes of the unit-norm operands (no range problem), a third of the
        # tensor-core work, |dS| = 1.3e-5 rms -> Q within 3e-3 rms (1.7e-2 max): a "bf16-grade" fast mode,
        # measured in tests/test_gpu_fullsize.py::test_score_gemm_precision_modes_at_config_eps
        self.proto_f16 = bool(proto_f16)
        dev = w_proj.device
        self.g_proj = torch.zeros_like(w_proj)
        # prototype gradients and the loss share one flat buffer [g_proto | g_bias | pad | loss]: one fill per step,
        # and - on several GPUs - one all-reduce that can start before the projection-weight gradient is folded
        kc, k = w_proto.numel(), b_proto.numel()
        off_b = (kc + 3) // 4 * 4
        off_l = off_b + (k + 3) // 4 * 4
        self.g_flat = torch.zeros(off_l + 4, dtype=torch.float32, device=dev)
        self.g_proto = self.g_flat[:kc].view_as(w_proto)
        self.g_bias = self.g_flat[off_b:off_b + k]
        self.loss_slot = self.g_flat[off_l:off_l + 1]
        self.loss_ring = torch.zeros(16, dtype=torch.float32, device=dev)     # losses of the last 16 steps
        self.m_proj = torch.zeros_like(w_proj)
        self.m_proto = torch.zeros_like(w_proto)
        self.m_bias = torch.zeros_like(b_proto)
        self.norms = L.larc_scratch(dev)
        self.steps = 0
        self.planes_ready = False

    @property
    def c(self):
        return self.w_proj.shape[0]

    @property
    def d(self):
        return self.w_proj.shape[1]

    @property
    def k(self):
        return self.w_proto.shape[0]

    def refresh_planes(self, need_bwd=True):
        want_lo = self.passes_fwd == 3
        self.wp_hi, self.wp_lo = L.split_planes(self.w_proj, want_lo=want_lo)
        if self.proto_f16:
            self.wk_f16 = L.round_f16(self.w_proto)
        else:
            self.wk_hi, self.wk_lo = L.split_planes(self.w_proto, want_lo=want_lo)
        if need_bwd:
            self.wkT_hi, self.wkT_lo = L.split_planes(self.w_proto, transpose=True, want_lo=self.passes_bwd == 3)
        self.planes_ready = True

    def zero_grad(self):
        self.g_proj.zero_()
        self.g_flat.zero_()

    def optimizer_step(self):
        first = 1 if self.steps == 0 else 0
        for p, g, m in ((self.w_proj, self.g_proj, self.m_proj), (self.w_proto, self.g_proto, self.m_proto),
                        (self.b_proto, self.g_bias, self.m_bias)):
            L.larc_sgd_(p, g, m, self.lr, self.momentum, self.trust, self.weight_decay, 1e-8, first, self.norms)
        self.steps += 1
        self.planes_ready = False


# ----------------------------------------------------------------------------------------
# stages
# ----------------------------------------------------------------------------------------

LOG2E = 1.4426950408889634


def _proto_scores(head: SwavHead, zn_hi, zn_lo, n, eps):
    """S = Zn Wk^T + b on the tensor cores; with `eps` the epilogue also accumulates the first
    Sinkhorn marginal u_k = sum_n exp(S_nk / eps), saving one pass over S."""
    u0 = None
    colexp = None
    if eps is not None:
        u0 = torch.zeros(head.k, dtype=torch.float32, device=zn_hi.device)
        colexp = (u0, LOG2E / eps)
    if head.proto_f16:     # zn_hi is the fp16 plane here (see scores_forward*)
        # CTA pairs (cta_group::2): M=256 MMAs over two SMs, each loading half of the Wk tile, with two
        # accumulator stages per SM so the (store-heavy) epilogue overlaps the next tile's MMAs
        s = L.gemm(zn_hi, None, head.wk_f16, None, n, head.k, head.c, 1, bias=head.b_proto,
                   tag="gemm_prototype_fwd", colexp=colexp, pair=True)
    else:
        s = L.gemm(zn_hi, zn_lo if head.passes_fwd == 3 else None, head.wk_hi, head.wk_lo, n, head.k, head.c,
                   head.passes_fwd, bias=head.b_proto, tag="gemm_prototype_fwd", colexp=colexp, pair=True)
    return s, u0


def _normalise(head: SwavHead, z, row_idx=None):
    """(zn_hi, zn_lo, inv_norm, operand of the score GEMM)"""
    lo = head.passes_fwd == 3 and not head.proto_f16
    if head.proto_f16:
        zn_hi, zn_lo, inv, zf = L.l2norm_split(z, want_lo=True, row_idx=row_idx, want_f16=True)
        return zn_hi, zn_lo, inv, zf
    zn_hi, zn_lo, inv = L.l2norm_split(z, want_lo=lo or head.passes_bwd == 3, row_idx=row_idx)
    return zn_hi, zn_lo, inv, zn_hi


def proj_activation(z, slope, out=None):
    """LeakyReLU(slope) of the '1-layer' projection network (ref :250-256), element-wise on Z"""
    return L.fused_bias_act_raw(z, None, None, 3, 0, slope, 1.0, out=out)


def proj_activation_bwd(dz, z_act, slope):
    """dZ * (z > 0 ? 1 : slope); z_act is the activation's OUTPUT (same sign as its input, like the reference's
    in-place LeakyReLU)"""
    return L.fused_bias_act_raw(dz, None, z_act, 3, 1, slope, 1.0)


def scores_forward(head: SwavHead, feats, out_h, out_w, hlen, row_img, row_src, nrows, eps=None):
    """gather -> projection -> normalise -> prototype scores.  Returns a dict of the
    tensors the backward needs."""
    if hlen % 8:
        raise ValueError("hlen must be a multiple of 8 (16-byte TMA row pitch of the bf16 operand planes)")
    lo = head.passes_fwd == 3
    a_hi, a_lo, _ = L.gather_rows(feats, out_h, out_w, hlen, row_img, row_src, nrows, want_lo=lo)
    z = L.gemm(a_hi, a_lo, head.wp_hi, head.wp_lo, nrows, head.c, hlen, head.passes_fwd, tag="gemm_projection_fwd")
    if head.proj_slope is not None:
        z = proj_activation(z, head.proj_slope)
    zn_hi, zn_lo, inv, za = _normalise(head, z)
    s, u0 = _proto_scores(head, za, zn_lo, nrows, eps)
    return dict(a_hi=a_hi, a_lo=a_lo, zn_hi=zn_hi, zn_lo=zn_lo, inv=inv, s=s, n=nrows, u0=u0,
                z_act=z if head.proj_slope is not None else None)


def sinkhorn_log_a(s, niters, eps, ws, n_total, group=None, r=None, c=None, pass_fn=None, log_a_fn=None,
                   u_first=None, cache16=False):
    """Sinkhorn-Knopp in scaling-vector form (ref :509-544) for ONE score matrix: niters streaming passes over
    the LOCAL rows of S; only u[K] crosses ranks (as in SwAV's distributed Sinkhorn), and c_n = 1/n_total uses
    the GLOBAL row count.  Returns log a[K]; Q = softmax_k(S/eps + log a).

    With `pass_fn` / `log_a_fn` (the CPU tests of the multi-rank logic inject torch stand-ins with the contract
    `pass_fn(s, inv_eps, first, u_in, r, c, n_total, ws) -> local column sums u[K]`, `log_a_fn(u, r)`) the
    exchange is a torch.distributed all-reduce; the CUDA path is `sinkhorn_multi`."""
    if pass_fn is None and log_a_fn is None:
        return sinkhorn_multi([dict(s=s, r=r, c=c, u_first=u_first)], niters, eps, ws, n_total, group,
                              cache16=cache16)[0]
    inv_eps = 1.0 / eps
    u = None
    for it in range(niters):
        if it == 0 and u_first is not None:
            u = u_first
        else:
            u = pass_fn(s, inv_eps, it == 0, u, r, c, n_total, ws)
        if group is not None:
            torch.distributed.all_reduce(u, group=group.pg)
    return log_a_fn(u, r)


# The iteration whose pass writes the 16-bit cache.  Not the first row-normalising one (1): the column scalings still move
# by four to five orders of magnitude after it (rho = a / a1 up to 1e5 at eps = 0.005), which lifts terms that underflowed
# the row-normalised fp16 plane back into play - with sharp score rows (a trained head: one prototype 30-50 nats above
# the rest) that costs up to 1.3e-2 in the codes.  Written one iteration later the plane keeps every case tried within
# 7e-4 (CPU test `test_sinkhorn_cached16_write_iteration`), for one more fp32 pass per call.
CACHE16_WRITE_IT = 2


def sinkhorn_multi(problems, niters, eps, ws, n_total, group=None, cache16=False):
    """Sinkhorn-Knopp (ref :509-544) on several independent score matrices (the s and t views of a patch):
    `problems` = [dict(s=S [n,K], r=, c=, u_first=)], returns [log a] per problem.

    One rank: the chains run one after the other, consecutive passes of a chain sweeping S in opposite directions
    (the tail of S that the L2 still holds is re-used).  Several ranks: the chains are interleaved pass by pass and
    the K-vector of column marginals goes through the low-latency NVLink exchange (`L.LLExchange`): the reduce
    kernel of chain A pushes its sums into every peer's buffer, chain B's pass runs meanwhile, and the next pass of
    chain A picks the peers' values up in its prologue - no collective call, no exposed latency, and a whole
    pass of slack against rank-to-rank jitter.  Without an exchange object the fallback is an NCCL all-reduce.

    `cache16`: the pass of iteration CACHE16_WRITE_IT (2) also stores its row-normalised terms as a 16-bit plane and
    the later passes stream that plane instead of S (`gx_sinkhorn_pass_cached`: half the bytes per pass, no exponentials).
    Only the column scalings log a come out of these passes - the codes are always evaluated from the fp32 scores -
    and they move by <~ 5e-4 relative (the training step's default; the API calls keep the fp32 passes).  Problems with
    non-uniform marginals (`r` / `c` given: source_pdf == 'image') always take the fp32 passes."""
    inv_eps = 1.0 / eps
    ll = group.ll if group is not None else None
    k = problems[0]["s"].shape[1]
    dev = problems[0]["s"].device
    if ll is not None and len(problems) > ll.channels:
        raise ValueError("more Sinkhorn chains than exchange channels")
    us = [None] * len(problems)

    def advance(ch, it):
        pb = problems[ch]
        if it == 0 and pb.get("u_first") is not None:
            parts, nparts = pb["u_first"], 1       # local u_k = sum_n exp(S_nk/eps) from the score GEMM's epilogue
        else:
            u_ll = ll.last(ch) if (ll is not None and it > 0) else None
            # (uniform marginals only: with source_pdf == 'image' an empty histogram bin gives a prototype a target
            #  mass of 1e-9 counts against ~10 - its column of the row-normalised plane is below the fp16 range)
            if cache16 and it >= CACHE16_WRITE_IT and niters > CACHE16_WRITE_IT + 1 and pb.get("r") is None and \
                    pb.get("c") is None:
                nparts = L.sinkhorn_pass_cached_parts(pb["s"], inv_eps, None if u_ll is not None else us[ch],
                                                      pb.get("r"), pb.get("c"), n_total, ws,
                                                      ws.cache16(ch, pb["s"].shape[0]), it == CACHE16_WRITE_IT,
                                                      u_ll=u_ll, reverse=(it & 1) == 1)
            else:
                nparts = L.sinkhorn_pass_parts(pb["s"], inv_eps, it == 0, None if u_ll is not None else us[ch],
                                               pb.get("r"), pb.get("c"), n_total, ws, u_ll=u_ll,
                                               reverse=(it & 1) == 1)
            parts = ws.partials
        if ll is not None:
            L.sinkhorn_reduce_send(parts, nparts, k, ll.next_send(ch))
            return
        if nparts == 1 and parts is pb.get("u_first"):
            us[ch] = parts
        else:
            if us[ch] is None or us[ch] is pb.get("u_first"):
                us[ch] = torch.empty((k,), dtype=torch.float32, device=dev)
            L.sinkhorn_reduce(parts, nparts, k, us[ch])
        if group is not None:
            torch.distributed.all_reduce(us[ch], group=group.pg)

    if group is None:
        for ch in range(len(problems)):
            for it in range(niters):
                advance(ch, it)
    else:
        for it in range(niters):
            for ch in range(len(problems)):
                advance(ch, it)
    if ll is not None:
        return [L.sinkhorn_log_a(None, pb.get("r"), u_ll=ll.last(ch), k=k, device=dev)
                for ch, pb in enumerate(problems)]
    return [L.sinkhorn_log_a(us[ch], pb.get("r")) for ch, pb in enumerate(problems)]


def scores_forward_dedup(head: SwavHead, z_all, row_idx, eps=None):
    """Per-patch part of the forward when every pixel has been projected once: gather +
    normalise the patch's rows of Z, prototype scores."""
    lo = head.passes_fwd == 3
    n = row_idx.numel()
    zn_hi, zn_lo, inv, za = _normalise(head, z_all, row_idx)
    s, u0 = _proto_scores(head, za, zn_lo, n, eps)
    return dict(zn_hi=zn_hi, zn_lo=zn_lo, inv=inv, s=s, n=n, u0=u0)


def scores_backward(head: SwavHead, fw, ds_hi, ds_lo, dz_rows_out=None):
    """Accumulates gWk += dS^T Zn and gWp += dZ^T A for one view-patch.  With `dz_rows_out`
    (fp32 [n,c] slice) the projection-weight gradient is deferred: dZ rows are stored and later
    folded per pixel (`project_backward_dedup`)."""
    n, k, c, d = fw["n"], head.k, head.c, head.d
    pb = head.passes_bwd
    dzn = L.gemm(ds_hi, ds_lo if pb == 3 else None, head.wkT_hi, head.wkT_lo if pb == 3 else None, n, c, k, pb,
                 tag="gemm_dzn_bwd", pair=True)
    kit = (n + 63) // 64
    sms = L.load().gx_sinkhorn_max_parts()
    bm = 256 if pb == 1 else 128     # the engine uses 256-row CTA tiles for single-pass GEMMs
    # CTA pairs: one unit of work = 256 x 256 outputs on two SMs (256 x 512 for single-pass GEMMs with c % 512 == 0)
    wide = 512 if (pb == 1 and c % 512 == 0 and n >= 2048) else 256
    sk1 = pick_split_k(math.ceil(k / 256) * math.ceil(c / wide), kit, sms // 2)
    L.gemm(ds_hi, ds_lo if pb == 3 else None, fw["zn_hi"], fw["zn_lo"] if pb == 3 else None, k, c, n, pb,
           out=head.g_proto, a_mn=True, b_mn=True, split_k=sk1, accumulate=True, tag="gemm_gproto_bwd", pair=True)
    if dz_rows_out is not None:
        if dz_rows_out.dtype == torch.bfloat16:      # bf16 backward: the rows are kept as one bf16 plane
            L.l2norm_bwd_split(dzn, fw["zn_hi"], fw["zn_lo"], fw["inv"], want_lo=False, out_hi=dz_rows_out)
        else:
            L.l2norm_bwd_split(dzn, fw["zn_hi"], fw["zn_lo"], fw["inv"], want_planes=False, out_f32=dz_rows_out)
        return
    if head.proj_slope is not None:
        dz = torch.empty((n, c), dtype=torch.float32, device=dzn.device)
        L.l2norm_bwd_split(dzn, fw["zn_hi"], fw["zn_lo"], fw["inv"], want_planes=False, out_f32=dz)
        dz_hi, dz_lo = L.split_planes(proj_activation_bwd(dz, fw["z_act"], head.proj_slope), want_lo=pb == 3)
    else:
        dz_hi, dz_lo = L.l2norm_bwd_split(dzn, fw["zn_hi"], fw["zn_lo"], fw["inv"], want_lo=pb == 3)
    sk2 = pick_split_k(math.ceil(c / bm) * math.ceil(d / 256), kit, sms)
    L.gemm(dz_hi, dz_lo, fw["a_hi"], fw["a_lo"] if pb == 3 else None, c, d, n, pb, out=head.g_proj, a_mn=True,
           b_mn=True, split_k=sk2, accumulate=True, tag="gemm_gproj_bwd")


def resolution_groups(feats, hlen):
    """Consecutive feature maps of equal resolution, cut at `hlen` channels of the concatenated
    per-pixel vector (ref :108-130): [dict(maps, h, w, off, keep)]."""
    groups, off = [], 0
    for f in feats:
        if off >= hlen:
            break
        h, w, c = f.shape[1], f.shape[2], f.shape[3]
        keep = min(c, hlen - off)
        if groups and groups[-1]["h"] == h and groups[-1]["w"] == w:
            groups[-1]["maps"].append(f)
            groups[-1]["keep"] += keep
        else:
            groups.append(dict(maps=[f], h=h, w=w, off=off, keep=keep))
        off += keep
    return groups


def project_all_pixels(wp_hi, wp_lo, feats, batch, out_h, out_w, hlen, passes, want_hi_only_planes=False, out=None,
                       out_planes=None, bilinear=False, labels=None, act_slope=None):
    """Z[pixel] = Wp . (nearest-upsampled, concatenated feature vector of the pixel) for EVERY
    pixel of `batch` images.  Upsampling and projection are both linear, so
    Z = sum_r upsample(F_r Wp[:, cols_r]^T): each resolution is projected at its native size
    (11x fewer flops than projecting the upsampled vectors for the ffhq pyramid) and the
    partial maps are summed per pixel.  Returns (Z fp32 [batch*H*W, C], levels) where each
    level keeps the bf16 planes of F_r for the weight-gradient GEMM."""
    c = wp_hi.shape[0]
    levels, parts = [], []
    for g in resolution_groups(feats, hlen):
        n = batch * g["h"] * g["w"]
        a_hi = torch.empty((n, g["keep"]), dtype=torch.bfloat16, device=wp_hi.device)
        a_lo = torch.empty_like(a_hi) if passes == 3 else None
        col = 0
        for f in g["maps"]:                       # K-concatenate the maps of this resolution
            cw = min(f.shape[3], g["keep"] - col)
            L.split_planes(f.view(n, f.shape[3])[:, :cw],
                           out=(a_hi[:, col:col + cw], a_lo[:, col:col + cw] if a_lo is not None else None))
            col += cw
        sl = slice(g["off"], g["off"] + g["keep"])
        p = L.gemm(a_hi, a_lo, wp_hi[:, sl], wp_lo[:, sl] if wp_lo is not None else None, n, c, g["keep"], passes,
                   tag="gemm_projection_fwd", pair=True)
        parts.append(p.view(batch, g["h"], g["w"], c))
        levels.append(dict(a_hi=a_hi, a_lo=None if want_hi_only_planes else a_lo, h=g["h"], w=g["w"], off=g["off"],
                           keep=g["keep"]))
    if len(parts) == 1 and parts[0].shape[1] == out_h and parts[0].shape[2] == out_w and out is None \
            and out_planes is None and labels is None:
        z = parts[0].view(-1, c)
        return (proj_activation(z, act_slope) if act_slope is not None else z), levels
    # nearest or bilinear (hf_interp): both are linear, so the per-resolution identity holds for either; `labels`:
    # the arg-max label map of predict_swav_codes, taken by the same kernel from the sums in registers
    if act_slope is not None:
        # '1-layer' projection network: the sums, then LeakyReLU; planes and labels come from the activated codes
        z = proj_activation(L.upsample_sum(parts, batch, out_h, out_w, bilinear=bilinear), act_slope, out=out)
        if out_planes is not None:
            L.split_planes(z, out=out_planes)
        if labels is not None:
            L.argmax_rows(z, out=labels)
        return z, levels
    z = L.upsample_sum(parts, batch, out_h, out_w, out=out, planes=out_planes, bilinear=bilinear, labels=labels)
    return z, levels


def project_backward_dedup(head: SwavHead, dz_rows, order, seg_off, levels, batch, out_h, out_w, bilinear=False,
                           z_act=None):
    """gWp[:, cols_r] += pool_r(dZ_pix)^T F_r per resolution, with dZ_pix[pixel] = sum of the dZ rows
    of all samples of that pixel and pool_r = block sums (the adjoint of nearest upsampling).
    z_act: the activated codes of every pixel ('1-layer' projection network) - the folded dZ of a pixel is
    multiplied by the LeakyReLU derivative there (the mask depends on the pixel only, so it commutes with the fold)."""
    pb = head.passes_bwd
    c = head.c
    npix = batch * out_h * out_w
    need_f32 = any(lv["h"] != out_h or lv["w"] != out_w for lv in levels)
    need_planes = any(lv["h"] == out_h and lv["w"] == out_w for lv in levels)
    if z_act is not None:
        _, _, f32 = L.segment_sum_rows(dz_rows, order, seg_off, npix, want_planes=False, want_f32=True)
        f32 = proj_activation_bwd(f32, z_act, head.proj_slope)
        hi, lo = L.split_planes(f32, want_lo=pb == 3) if need_planes else (None, None)
    else:
        hi, lo, f32 = L.segment_sum_rows(dz_rows, order, seg_off, npix, want_lo=pb == 3, want_planes=need_planes,
                                         want_f32=need_f32)
    cur = dict(h=out_h, w=out_w, f32=f32.view(batch, out_h, out_w, c) if f32 is not None else None, hi=hi, lo=lo)
    bm = 256 if pb == 1 else 128
    sms = L.load().gx_sinkhorn_max_parts()
    order_lv = sorted(levels, key=lambda lv: -lv["h"] * lv["w"])
    full = cur
    for i, lv in enumerate(order_lv):
        if bilinear and (lv["h"] != out_h or lv["w"] != out_w):
            if lv["h"] != cur["h"] or lv["w"] != cur["w"] or cur is full:
                # bilinear upsampling from different resolutions does not compose: every level folds the
                # full-resolution dZ with its own adjoint (x then y)
                f = L.pool_bilinear_adjoint(full["f32"], lv["h"], lv["w"])
                hi, lo = L.split_planes(f.view(-1, c), want_lo=pb == 3)
                cur = dict(h=lv["h"], w=lv["w"], f32=None, hi=hi, lo=lo)
        elif lv["h"] != cur["h"] or lv["w"] != cur["w"]:
            more = any(o["h"] * o["w"] < lv["h"] * lv["w"] for o in order_lv[i + 1:])
            f, hi, lo = L.pool_sum(cur["f32"], lv["h"], lv["w"], want_f32=more, want_planes=True, want_lo=pb == 3)
            cur = dict(h=lv["h"], w=lv["w"], f32=f, hi=hi, lo=lo)
        n = batch * lv["h"] * lv["w"]
        sk = pick_split_k(math.ceil(c / bm) * math.ceil(lv["keep"] / 256), (n + 63) // 64, sms)
        L.gemm(cur["hi"], cur["lo"], lv["a_hi"], lv["a_lo"] if pb == 3 else None, c, lv["keep"], n, pb,
               out=head.g_proj[:, lv["off"]:lv["off"] + lv["keep"]], a_mn=True, b_mn=True, split_k=sk, accumulate=True,
               tag="gemm_gproj_bwd")


@dataclass
class DistGroup:
    pg: object
    rank: int
    world: int
    ll: Optional[object] = None          # L.LLExchange: NVLink exchange of the Sinkhorn marginals (CUDA ranks)
    ll_failed: bool = False              # set when the exchange could not be created (NCCL is used instead)

    def ensure_ll(self, k, device):
        """create the low-latency exchange for K prototypes (collective: every rank must call it).  If any rank cannot
        set it up (CUDA IPC unavailable between the processes, e.g. separate containers), ALL ranks fall back to the
        NCCL all-reduce of the marginals - decided together, so the ranks never disagree on the transport."""
        if self.ll is not None and self.ll.k == k:
            return self.ll
        if self.ll is not None:
            self.ll.close()
            self.ll = None
        import warnings
        ll, err = None, None
        try:
            ll = L.LLExchange(self.pg, self.rank, self.world, k, device)
        except Exception as e:          # noqa: BLE001 - any failure means "no peer memory here"
            err = e
        ok = torch.tensor([1 if ll is not None else 0], dtype=torch.int32, device=device)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=self.pg)
        if int(ok.item()) == 1:
            self.ll = ll
        else:
            if ll is not None:
                ll.close()
            if self.rank == 0:
                warnings.warn(f"ganecdotes_b200: NVLink exchange unavailable ({err!r}); the Sinkhorn marginals use the "
                              f"NCCL all-reduce")
            self.ll = None
            self.ll_failed = True
        return self.ll


def image_marginals(feats, out_h, out_w, hlen, index_map, k, n):
    """source_pdf == 'image' (ref :361-362, :523-532): per-pixel L2 norm of the transformed
    feature tensor -> histograms with K and N bins -> Sinkhorn marginals r[K], c[N].
    index_map: int32 [H*W] device tensor (rotate/flip source pixels).  Single image only."""
    _, _, _, nrm = L.gather_rows(feats, out_h, out_w, hlen, torch.zeros_like(index_map), index_map,
                                 index_map.numel(), want_lo=False, want_planes=False, want_norm=True)
    img = nrm.view(1, out_h, out_w)
    histb = torch.histc(img, n) + 1e-9
    histb[0] = histb[1]
    histb = histb / histb.sum()
    histk = torch.histc(img, k) + 1e-9
    histk[0] = histk[1]
    histk = histk / histk.sum()
    return histk.contiguous(), histb.contiguous()


@dataclass
class StepConfig:
    hlen: int
    patch_size: Optional[int]
    num_patches: int
    niters: int
    eps: float
    temperature: float
    truncation: float
    perturb_std: List[float]
    need_image: bool = False
    source_pdf: str = 'uniform'
    dedup: Optional[bool] = None   # project every pixel once (None: automatic, when P*N > H*W)
    hf_interp: str = 'nearest'     # 'nearest' | 'bilinear' (ref :112-126); bilinear needs the all-pixel path
    sinkhorn_cache16: Optional[bool] = None   # Sinkhorn passes 3.. stream a 16-bit cache of the scaled kernel matrix
    #                                           (sinkhorn_multi); None: on, unless GX_SINKHORN_CACHE16=0


@dataclass
class StepInputs:
    """Device-resident inputs of one optimiser step (what `prepare_step_inputs` uploads)."""
    z: torch.Tensor                    # [B, D]  (view of zcat)
    views: dict                        # name -> (layer_no list, pert rows [2B, D] on device (view of zcat))
    rows: dict                         # name -> (row_src [P, B*N] int32, row_img [B*N] int32)
    h2d_bytes: int = 0
    index_maps: Optional[dict] = None  # name -> int32 [H*W] (source_pdf == 'image', single latent)
    dedup: Optional[dict] = None       # name -> (row_idx [P, B*N] int32, order int32, seg_off int32 [B*H*W+1]);
    #                                    {} = "all-pixel path, segments not built yet" (built on the device by the step)
    ready: Optional[object] = None     # CUDA event recorded on the upload stream (None: same stream)
    zcat: Optional[torch.Tensor] = None      # [B + 2B + 2B, D]: latents, view-s draws, view-t draws (one upload)
    layer_no: Optional[torch.Tensor] = None  # int32 [2B]: perturbed layer per (view, latent)
    sigma: Optional[torch.Tensor] = None     # fp32 [2B]: perturb_std of that layer
    rows_both: Optional[torch.Tensor] = None  # int32 [2P, B*N]: row_src of view s then view t (one upload)


def use_dedup(cfg: StepConfig, out_h, out_w) -> bool:
    """The projection depends on the pixel, not on the patch, and the all-pixel projection runs
    at each level's native resolution (`project_all_pixels`, ~11x cheaper per pixel than projecting
    gathered rows for the ffhq pyramid): project every pixel once and let the patches gather rows
    of Z unless the patches touch only a small fraction of the image."""
    if getattr(cfg, "hf_interp", "nearest") == 'bilinear':
        if cfg.dedup is False:
            raise ValueError("hf_interp='bilinear' runs on the all-pixel projection path (dedup=False gathers "
                             "nearest-upsampled rows)")
        return True
    if cfg.dedup is not None:
        return bool(cfg.dedup)
    n = cfg.patch_size if cfg.patch_size is not None else out_h * out_w
    return 4 * cfg.num_patches * n > out_h * out_w


def prepare_step_inputs(gen, draws: StepDraws, cfg: StepConfig, device, stream=None) -> StepInputs:
    """Host bookkeeping + host->device copies of one step: latents, the two perturbation
    draws per latent-view that are actually used, and the sampled-pixel source indices.
    With `stream` (a side CUDA stream) the uploads and the index bookkeeping kernels are issued
    there, so the inputs of step i+1 are staged while the compute stream runs step i; the step
    waits on `ready`."""
    if stream is not None:
        main = torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            inp = prepare_step_inputs(gen, draws, cfg, device)
            inp.ready = torch.cuda.Event()
            inp.ready.record(stream)
        for t in _tensors_of(inp):      # allocated on the side stream, consumed on the compute stream
            t.record_stream(main)
        return inp
    out_h = out_w = gen.size
    views, rows, nbytes = {}, {}, 0
    b = draws.z.shape[0]
    prs, lno, sig = [], [], []
    for name, view in (("s", draws.view_s), ("t", draws.view_t)):
        for i, l in enumerate(view.layer_no):
            prs += [view.pert_z[i, 2 * l], view.pert_z[i, 2 * l + 1]]
            lno.append(int(l))
            sig.append(float(cfg.perturb_std[l]))
    zcat = _upload(torch.cat([draws.z.float(), torch.stack(prs).float()]), device)       # one copy: [5B, D]
    layer_no = _upload(torch.tensor(lno, dtype=torch.int32), device)
    sigma = _upload(torch.tensor(sig, dtype=torch.float32), device)
    nbytes += zcat.numel() * 4 + 8 * len(lno)
    host_rows = []
    for vi, (name, view) in enumerate((("s", draws.view_s), ("t", draws.view_t))):
        views[name] = (list(view.layer_no), zcat[b + 2 * b * vi: b + 2 * b * (vi + 1)])
        host_rows.append(build_row_indices(out_h, out_w, view, draws.perms, cfg.patch_size, "cpu"))
    # one upload for both views: [2P, B*N] source pixels (view s patches first), one image-of-sample vector
    npatch = host_rows[0][0].shape[0]
    rs_both = _upload(torch.cat([host_rows[0][0], host_rows[1][0]]), device)
    ri = _upload(host_rows[0][1], device)
    nbytes += rs_both.numel() * 4 + ri.numel() * 4
    for vi, name in enumerate(("s", "t")):
        rows[name] = (rs_both[vi * npatch:(vi + 1) * npatch], ri)
    dedup = {} if use_dedup(cfg, out_h, out_w) else None
    index_maps = None
    if cfg.source_pdf == 'image':
        if draws.z.shape[0] != 1:
            raise NotImplementedError("source_pdf='image' is defined for one latent per step (as in the reference)")
        index_maps = {}
        for name, view in (("s", draws.view_s), ("t", draws.view_t)):
            m = rotate_flip_index_map(out_h, out_w, view.angle[0], view.flip[0]).to(torch.int32)
            index_maps[name] = _upload(m, device)
            nbytes += m.numel() * 4
    return StepInputs(z=zcat[:b], views=views, rows=rows, h2d_bytes=nbytes, index_maps=index_maps, dedup=dedup,
                      zcat=zcat, layer_no=layer_no, sigma=sigma, rows_both=rs_both)


def build_pixel_segments(inp: StepInputs, batch, hw):
    """Device bookkeeping of the all-pixel path (part of the step): for each view the row of Z every sample reads
    and the CSR list of the samples of every pixel (`gx_pixel_segments`: counting sort, deterministic order)."""
    if inp.dedup is None or inp.dedup:
        return
    # both views in one call: the views' images are b apart in the 2b-image batch the synthesis runs on
    npatch = inp.rows_both.shape[0] // 2
    ridx, order, seg_off = L.pixel_segments(inp.rows_both, inp.rows["s"][1], hw, 2 * batch * hw, npatch, batch)
    inp.dedup.update(ridx=ridx, order=order, seg_off=seg_off, npatch=npatch)


def _tensors_of(inp: StepInputs):
    out = [inp.zcat, inp.layer_no, inp.sigma]
    out.append(inp.rows["s"][1])
    for d in (inp.index_maps, ):
        if d:
            out += list(d.values())
    out.append(inp.rows_both)
    return [t for t in out if t is not None and t.is_cuda]


@torch.no_grad()
def swav_train_step_device(gen, head: SwavHead, mean_latent, inp: StepInputs, cfg: StepConfig,
                           group: Optional[DistGroup] = None, ws: Optional[L.SinkhornWorkspace] = None):
    """One optimiser step on this rank's latents from device-resident inputs.  Returns the
    (global) loss as a 0-dim device tensor; no host synchronisation."""
    dev = head.w_proj.device
    b = inp.z.shape[0]
    world = group.world if group is not None else 1
    if inp.ready is not None:
        torch.cuda.current_stream().wait_event(inp.ready)
    # prototype re-normalisation every step (ref :328-331), then operand planes
    L.normalize_rows_(head.w_proto)
    head.refresh_planes()
    head.zero_grad()
    if ws is None:      # kept on the head: the workspace owns the 16-bit Sinkhorn cache (2 x N x K x 2 bytes), not
        ws = getattr(head, "_sk_ws", None)   # something to allocate per step
        if ws is None or ws.k != head.k or ws.partials.device != dev:
            ws = head._sk_ws = L.SinkhornWorkspace(head.k, dev)
    # one pass of the mapping network over the latents and both views' perturbation draws
    styled = gen.style(inp.zcat)
    feats = {}
    out_h = out_w = gen.size
    build_pixel_segments(inp, b, out_h * out_w)
    dedup = inp.dedup is not None
    if cfg.hlen % 8:
        raise ValueError("hlen must be a multiple of 8 (16-byte TMA row pitch of the bf16 operand planes)")
    n_patch_rows = b * (cfg.patch_size if cfg.patch_size is not None else out_h * out_w)
    # both views go through the synthesis network as ONE batch of 2b images (they differ only in W+):
    # half the launches, and the latency-bound 4x4 ... 16x16 layers do twice the work per launch
    wplus = L.view_wplus(styled[:b], styled[b:], inp.layer_no, inp.sigma, mean_latent.reshape(-1).float().contiguous(),
                         cfg.truncation, gen.n_latent)
    _, f_both = gen.synthesize(wplus, None, need_image=cfg.need_image)
    for vi, name in enumerate(("s", "t")):
        feats[name] = [t[vi * b:(vi + 1) * b] for t in f_both]
    if dedup:
        # every pixel of BOTH views projected in one set of per-resolution GEMMs (2b images): half the launches of
        # the small levels; rows [0, b*hw) of Z belong to view s, the rest to view t
        z_all, levels = project_all_pixels(head.wp_hi, head.wp_lo, f_both, 2 * b, out_h, out_w, cfg.hlen,
                                           head.passes_fwd, want_hi_only_planes=head.passes_bwd != 3,
                                           bilinear=cfg.hf_interp == 'bilinear', act_slope=head.proj_slope)
        # dZ rows of every sample, folded per pixel after the last patch: fp32, or - with bf16 backward operands
        # (passes_bwd == 1, the default) - one bf16 plane: half the bytes written here and gathered by the segment sum
        dz_rows = torch.empty((2 * cfg.num_patches * n_patch_rows, head.c), device=dev,
                              dtype=torch.bfloat16 if head.passes_bwd == 1 else torch.float32)

    n_local = b * (cfg.patch_size if cfg.patch_size is not None else out_h * out_w)
    n_total = n_local * world
    grad_scale = 1.0 / (n_total * cfg.num_patches)
    maxp = L.load().gx_loss_max_parts()
    loss_parts = torch.zeros((cfg.num_patches, maxp), dtype=torch.float32, device=dev)
    if group is not None and group.ll is None and not group.ll_failed and \
            os.environ.get("GX_SINKHORN_EXCHANGE", "ll") != "nccl":
        group.ensure_ll(head.k, dev)           # NVLink exchange of the marginals (GX_SINKHORN_EXCHANGE=nccl: A/B)
    cache16 = cfg.sinkhorn_cache16
    if cache16 is None:
        cache16 = os.environ.get("GX_SINKHORN_CACHE16", "1") != "0"
    for p in range(cfg.num_patches):
        fw = {}
        for name in ("s", "t"):
            if dedup:
                vi = 0 if name == "s" else 1
                fw[name] = scores_forward_dedup(head, z_all, inp.dedup["ridx"][vi * cfg.num_patches + p], cfg.eps)
            else:
                row_src, row_img = inp.rows[name]
                fw[name] = scores_forward(head, feats[name], out_h, out_w, cfg.hlen, row_img, row_src[p], n_local,
                                          cfg.eps)
        rc_s = rc_t = (None, None)
        if cfg.source_pdf == 'image':
            rc_s = image_marginals(feats["s"], out_h, out_w, cfg.hlen, inp.index_maps["s"], head.k, n_local)
            rc_t = image_marginals(feats["t"], out_h, out_w, cfg.hlen, inp.index_maps["t"], head.k, n_local)
        la_s, la_t = sinkhorn_multi([dict(s=fw["s"]["s"], r=rc_s[0], c=rc_s[1], u_first=fw["s"]["u0"]),
                                     dict(s=fw["t"]["s"], r=rc_t[0], c=rc_t[1], u_first=fw["t"]["u0"])],
                                    cfg.niters, cfg.eps, ws, n_total, group, cache16=cache16)
        lo = head.passes_bwd == 3
        _, ds_s, ds_t, _, _ = L.swav_loss(fw["s"]["s"], fw["t"]["s"], 1.0 / cfg.eps, 1.0 / cfg.temperature,
                                          la_s, la_t, grad_scale, want_lo=lo, loss_parts=loss_parts[p],
                                          db_accum=head.g_bias)
        fw["s"].pop("s"), fw["t"].pop("s")
        for vi, (name, ds) in enumerate((("s", ds_s), ("t", ds_t))):
            r0 = (vi * cfg.num_patches + p) * n_local
            out = dz_rows[r0:r0 + n_local] if dedup else None
            scores_backward(head, fw[name], ds[0], ds[1], out)
    # loss = sum of the per-CTA partial sums of every patch / (N_global * P), into the tail of the flat gradient
    L.colsum(loss_parts, cfg.num_patches * maxp, 1, head.loss_slot, scale=grad_scale)
    pending = None
    if group is not None:
        # prototype gradients + loss travel while the projection-weight gradient is still being folded
        pending = torch.distributed.all_reduce(head.g_flat, group=group.pg, async_op=True)
    if dedup:
        project_backward_dedup(head, dz_rows, inp.dedup["order"], inp.dedup["seg_off"], levels, 2 * b, out_h, out_w,
                               bilinear=cfg.hf_interp == 'bilinear',
                               z_act=z_all if head.proj_slope is not None else None)
    if group is not None:
        torch.distributed.all_reduce(head.g_proj, group=group.pg)
        pending.wait()
    loss = head.loss_ring[head.steps % head.loss_ring.numel():][:1]
    L.colsum(head.loss_slot, 1, 1, loss)           # keep the step's loss while the next step re-uses the buffer
    head.optimizer_step()
    return loss.view(())


@torch.no_grad()
def swav_train_step(gen, head: SwavHead, mean_latent, draws: StepDraws, cfg: StepConfig,
                    group: Optional[DistGroup] = None, ws: Optional[L.SinkhornWorkspace] = None):
    """Public step: host draws in, loss tensor out (host bookkeeping + uploads + device step)."""
    inp = prepare_step_inputs(gen, draws, cfg, head.w_proj.device)
    return swav_train_step_device(gen, head, mean_latent, inp, cfg, group, ws)


@torch.no_grad()
def predict_codes(gen, w_proj, w, mean_latent, truncation, hlen, passes=3, images_per_chunk=16, want_planes=False,
                  hf_interp='nearest', proj_slope=None):
    """predict_swav_codes (ref :659-693): generator forward with the fixed noise buffers,
    per-pixel vectors, projection only, arg-max over the code channels.
    Returns (codes [B,C,H,W] fp32 in channels_last memory, labels int64 [B,H,W]); with want_planes also the
    bf16 (hi, lo) planes [B*H*W, C] of the codes (operand of the one-shot segmentor head), emitted by the
    same kernel that writes the codes."""
    if hlen % 8:
        raise ValueError("hlen must be a multiple of 8 (16-byte TMA row pitch of the bf16 operand planes)")
    dev = w_proj.device
    mean = mean_latent.reshape(-1).float().contiguous()
    w = w.to(dev).float().contiguous()
    wt = L.truncate(w, mean, truncation) if truncation < 1 else w
    latent = wt.unsqueeze(1).expand(-1, gen.n_latent, -1) if wt.dim() == 2 else wt      # broadcast W+: no copy
    _, feats = gen.synthesize(latent, None, need_image=False)
    b = latent.shape[0]
    h = wd = gen.size
    c = w_proj.shape[0]
    wp_hi, wp_lo = L.split_planes(w_proj.contiguous(), want_lo=passes == 3)
    z = torch.empty((b * h * wd, c), dtype=torch.float32, device=dev)
    labels = torch.empty((b * h * wd,), dtype=torch.int64, device=dev)
    z_hi = z_lo = None
    if want_planes:
        z_hi = torch.empty((b * h * wd, c), dtype=torch.bfloat16, device=dev)
        z_lo = torch.empty_like(z_hi)
    for i0 in range(0, b, images_per_chunk):
        i1 = min(b, i0 + images_per_chunk)
        sub = [f[i0:i1] for f in feats]
        zc = z[i0 * h * wd: i1 * h * wd]
        pl = (z_hi[i0 * h * wd: i1 * h * wd], z_lo[i0 * h * wd: i1 * h * wd]) if want_planes else None
        project_all_pixels(wp_hi, wp_lo, sub, i1 - i0, h, wd, hlen, passes, out=zc, out_planes=pl,
                           bilinear=hf_interp == 'bilinear', labels=labels[i0 * h * wd: i1 * h * wd],
                           act_slope=proj_slope)
    preds = z.view(b, h, wd, c).permute(0, 3, 1, 2)
    if want_planes:
        return preds, labels.view(b, h, wd), (z_hi, z_lo)
    return preds, labels.view(b, h, wd)


class TrainGraph:
    """One optimiser step (`swav_train_step_device`) captured as a CUDA graph and replayed on new inputs: ~360
    launches issued back to back instead of one stream launch each (the bubbles between launches are ~2 % of the
    step; the host side shrinks from ~7 ms of launch calls to one replay).

    The caller must have run at least one eager step with the same shapes before (lazy caches, function attributes,
    the momentum buffers' first-step branch).  Everything that is a host-side scalar at launch time is baked in: the
    learning rate (no `use_scheduler`), the step shapes, the patch count.  Inputs are copied into static buffers
    (five small device copies) before each replay; the loss comes back as a copy of the graph's loss slot."""

    def __init__(self, gen, head: SwavHead, mean_latent, inp: StepInputs, cfg: StepConfig,
                 group: Optional[DistGroup] = None, ws: Optional[L.SinkhornWorkspace] = None):
        if head.steps == 0:
            raise RuntimeError("TrainGraph: run one eager step first (the first optimiser step initialises the momentum "
                               "buffers on a different branch)")
        if cfg.source_pdf != 'uniform':
            raise NotImplementedError("TrainGraph: source_pdf='uniform' only (the image marginals use eager histograms)")
        self.gen, self.head, self.mean_latent, self.cfg, self.group = gen, head, mean_latent, cfg, group
        self.ws = ws or L.SinkhornWorkspace(head.k, head.w_proj.device)
        b = inp.z.shape[0]
        self.static = self._clone_inputs(inp, b)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        launches = L.launch_count
        steps_before = head.steps
        with torch.cuda.graph(self.graph):
            self.loss_slot = swav_train_step_device(gen, head, mean_latent, self.static, cfg, group, self.ws)
        head.steps = steps_before            # the capture did not execute anything
        head.planes_ready = True
        self.launches_per_replay = L.launch_count - launches

    @staticmethod
    def _clone_inputs(inp: StepInputs, b):
        zcat = inp.zcat.clone()
        rows_both = inp.rows_both.clone()
        ri = inp.rows["s"][1].clone()
        npatch = rows_both.shape[0] // 2
        rows = {name: (rows_both[vi * npatch:(vi + 1) * npatch], ri) for vi, name in enumerate(("s", "t"))}
        views = {name: (list(inp.views[name][0]), zcat[b + 2 * b * vi: b + 2 * b * (vi + 1)])
                 for vi, name in enumerate(("s", "t"))}
        return StepInputs(z=zcat[:b], views=views, rows=rows, h2d_bytes=inp.h2d_bytes, index_maps=None,
                          dedup={} if inp.dedup is not None else None, zcat=zcat, layer_no=inp.layer_no.clone(),
                          sigma=inp.sigma.clone(), rows_both=rows_both)

    @torch.no_grad()
    def __call__(self, inp: StepInputs):
        """replay on the inputs of a new step; returns the loss as a 0-dim device tensor"""
        if inp.ready is not None:
            torch.cuda.current_stream().wait_event(inp.ready)
        st = self.static
        st.zcat.copy_(inp.zcat, non_blocking=True)
        st.layer_no.copy_(inp.layer_no, non_blocking=True)
        st.sigma.copy_(inp.sigma, non_blocking=True)
        st.rows_both.copy_(inp.rows_both, non_blocking=True)
        st.rows["s"][1].copy_(inp.rows["s"][1], non_blocking=True)
        self.graph.replay()
        self.head.steps += 1
        L._count(self.launches_per_replay)
        return self.loss_slot.clone()


class PredictGraph:
    """`predict_codes` for a fixed batch size, captured once as a CUDA graph and replayed: the label-map path of a
    batch is ~150 launches of which many (the 4x4 ... 32x32 synthesis layers, the small per-resolution projections)
    are shorter than the gap between two stream launches; a graph replay issues them back to back.

        g = PredictGraph(gen, w_proj, mean_latent, truncation, hlen, batch=16)
        preds, labels = g(w)            # views of the graph's static outputs: valid until the next call

    The generator weights and `w_proj` are baked into the captured launches by ADDRESS (their values may change in
    place; re-create the object if the tensors are replaced).  Results are bit-identical to `predict_codes`."""

    def __init__(self, gen, w_proj, mean_latent, truncation, hlen, batch, passes=3, want_planes=False,
                 hf_interp='nearest'):
        dev = w_proj.device
        self.args = (gen, w_proj, mean_latent, truncation, hlen, passes, want_planes, hf_interp)
        self.w_static = torch.zeros((batch, gen.style_dim), dtype=torch.float32, device=dev)
        self.batch = batch
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up outside the capture: lazy caches, function attributes
            for _ in range(2):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        launches = L.launch_count
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        self.launches_per_replay = L.launch_count - launches

    def _run(self):
        gen, w_proj, mean_latent, truncation, hlen, passes, want_planes, hf_interp = self.args
        return predict_codes(gen, w_proj, self.w_static, mean_latent, truncation, hlen, passes, self.batch,
                             want_planes, hf_interp)

    @torch.no_grad()
    def __call__(self, w):
        if tuple(w.shape) != tuple(self.w_static.shape):
            raise ValueError(f"PredictGraph was captured for w of shape {tuple(self.w_static.shape)}")
        self.w_static.copy_(w, non_blocking=True)
        self.graph.replay()
        L._count(self.launches_per_replay)
        return self.out
