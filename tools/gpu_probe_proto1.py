import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
torch.manual_seed(0)
n, c, k = 160000, 512, 5000
z = torch.randn(n, c, device="cuda").to(torch.bfloat16)
w = torch.randn(k, c, device="cuda").to(torch.bfloat16)
b = torch.randn(k, device="cuda")
out = torch.empty(n, k, device="cuda")
for _ in range(3):
    L.gemm(z, z, w, w, n, k, c, 3, out=out, bias=b)
torch.cuda.synchronize()
print("done")
