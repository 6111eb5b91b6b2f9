"""The three dominant tcgen05 GEMMs of the ffhq-256 step exactly as the engine launches them (B = 8 latents:
N = 160000 rows, C = 512, K = 5000), for `ncu --set full -k regex:gx_umma_kernel -s 3 -c 3`:
  gemm_prototype_fwd  (bf16x3 split, CTA pairs, fused bias + first Sinkhorn marginal)
  gemm_dzn_bwd        (bf16x1, CTA pairs)          dZn = dS Wk
  gemm_gproto_bwd     (bf16x1, CTA pairs, MN-major operands, split-K, accumulate)   gWk += dS^T Zn
One warm-up round (3 launches) first, then the captured round."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganecdotes_b200 import _lib as L  # noqa: E402
from ganecdotes_b200.hfc_with_swav import engine as E  # noqa: E402


def main():
    torch.manual_seed(0)
    n, c, k, d = 160000, 512, 5000, 5376
    dev = "cuda"
    head = E.SwavHead((torch.randn(c, d) / d ** 0.5).to(dev), torch.nn.functional.normalize(torch.randn(k, c), dim=1).to(dev),
                      (0.01 * torch.randn(k)).to(dev), 0.01, 0.9, 0.01, 3, 1)
    head.refresh_planes()
    z = torch.randn(n, c, device=dev)
    ds_hi = (1e-6 * torch.randn(n, k, device=dev)).to(torch.bfloat16)
    dz_rows = torch.empty(n, c, device=dev)
    for _ in range(2):
        zn_hi, zn_lo, inv, za = E._normalise(head, z)
        s, u0 = E._proto_scores(head, za, zn_lo, n, 0.005)
        fw = dict(zn_hi=zn_hi, zn_lo=zn_lo, inv=inv, n=n)
        E.scores_backward(head, fw, ds_hi, None, dz_rows)
        torch.cuda.synchronize()
    print("ok", float(s[0, 0]), float(u0[0]))


if __name__ == "__main__":
    main()
