"""The dominant kernels of the ffhq-256 step exactly as the engine launches them (B = 8 latents: N = 160000 rows,
C = 512, K = 5000), two rounds of six launches matching `sinkhorn_pass|swav_loss|gx_umma_kernel`:
  gemm_prototype_fwd   bf16x3 split, CTA pairs, fused bias + first Sinkhorn marginal       S = Zn Wk^T + b
  sinkhorn_pass        the cache-writing iteration: fp32 scores in, 16-bit cache out (reverse sweep);
                       a later iteration: the pass over the cache (forward sweep)                 (serpentine passes)
  swav_loss_fwd_bwd    power-ratio kernel (T / eps = 2)
  gemm_dzn_bwd         bf16x1, 256 x 512 CTA-pair tiles                                     dZn = dS Wk
  gemm_gproto_bwd      bf16x1, 256 x 512 CTA-pair tiles, MN-major operands, split-K         gWk += dS^T Zn
Capture the second round:  ncu --set full -k regex:'sinkhorn_pass|swav_loss|gx_umma_kernel' -s 6 -c 6"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganecdotes_b200 import _lib as L  # noqa: E402
from ganecdotes_b200.hfc_with_swav import engine as E  # noqa: E402


def main():
    torch.manual_seed(0)
    n, c, k, d = 160000, 512, 5000, 5376
    eps, temp = 0.005, 0.01
    dev = "cuda"
    head = E.SwavHead((torch.randn(c, d) / d ** 0.5).to(dev),
                      torch.nn.functional.normalize(torch.randn(k, c), dim=1).to(dev),
                      (0.01 * torch.randn(k)).to(dev), 0.01, 0.9, 0.01, 3, 1)
    head.refresh_planes()
    z = torch.randn(n, c, device=dev)
    s_t = 0.05 * torch.randn(n, k, device=dev)
    ws = L.SinkhornWorkspace(k, dev)
    dz_rows = torch.empty(n, c, device=dev)
    for _ in range(2):
        zn_hi, zn_lo, inv, za = E._normalise(head, z)
        s, u0 = E._proto_scores(head, za, zn_lo, n, eps)                                   # gx_umma (1)
        u = torch.empty(k, device=dev)
        cache = ws.cache16(0, n)
        np_ = L.sinkhorn_pass_cached_parts(s, 1.0 / eps, u0, None, None, n, ws, cache, True, reverse=True)   # pass (2)
        L.sinkhorn_reduce(ws.partials, np_, k, u)
        np_ = L.sinkhorn_pass_cached_parts(s, 1.0 / eps, u, None, None, n, ws, cache, False, reverse=False)  # pass (3)
        L.sinkhorn_reduce(ws.partials, np_, k, u)
        la = L.sinkhorn_log_a(u, None)
        _, ds_s, ds_t, _, _ = L.swav_loss(s, s_t, 1.0 / eps, 1.0 / temp, la, la, 1e-6)      # loss (4)
        fw = dict(zn_hi=zn_hi, zn_lo=zn_lo, inv=inv, n=n)
        E.scores_backward(head, fw, ds_s[0], None, dz_rows)                                 # gx_umma (5), (6)
        torch.cuda.synchronize()
    print("ok", float(s[0, 0]), float(u[0]))


if __name__ == "__main__":
    main()
