"""One configuration of the score GEMM for an ncu capture: MODE = f16 (CTA pairs) | x3pair | x3 | x1"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
torch.manual_seed(0)
n, c, k = 160000, 512, 5000
z = torch.randn(n, c, device="cuda") / 22
w = torch.randn(k, c, device="cuda") / 22
b = torch.randn(k, device="cuda") * 0.01
out = torch.empty(n, k, device="cuda")
u = torch.zeros(k, device="cuda")
for _ in range(3):
    if mode == "f16":
        L.gemm(z.half(), None, w.half(), None, n, k, c, 1, out=out, bias=b, colexp=(u, 20.0), pair=True)
    elif mode == "x3pair":
        zb, wb = z.bfloat16(), w.bfloat16()
        L.gemm(zb, zb, wb, wb, n, k, c, 3, out=out, bias=b, colexp=(u, 20.0), pair=True)
    elif mode == "x1":
        L.gemm(z.bfloat16(), None, w.bfloat16(), None, n, k, c, 1, out=out, bias=b, force_m128=True)
    else:
        zb, wb = z.bfloat16(), w.bfloat16()
        L.gemm(zb, zb, wb, wb, n, k, c, 3, out=out, bias=b, colexp=(u, 20.0))
torch.cuda.synchronize()
print("done")
