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
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for passes in (3, 1):
    for bn, st in ((256, 0), (128, 0)):
        t = timeit(lambda: L.gemm(z, z if passes == 3 else None, w, w if passes == 3 else None, n, k, c, passes, out=out, bias=b, block_n=bn, stages=st))
        print(f"GX_UMMA_DEBUG={os.environ.get('GX_UMMA_DEBUG','0')} proto-shaped gemm passes={passes} bn={bn}: {t:.3f} ms  {2*n*k*c/t/1e9:.0f} TF alg", flush=True)
