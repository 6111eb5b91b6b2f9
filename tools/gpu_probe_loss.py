import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
torch.manual_seed(0)
n, k = 160000, 5000
s = (0.05 * torch.randn(n, k, device="cuda")); t = (0.05 * torch.randn(n, k, device="cuda"))
la = torch.zeros(k, device="cuda")
def run(): return L.swav_loss(s, t, 200.0, 100.0, la, la, 1.0 / n)
out = run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"GX_LOSS_MODE={os.environ.get('GX_LOSS_MODE','0')}: {ms:.3f} ms  {n*k*12/ms/1e6:.0f} GB/s  loss={out[0].sum().item()/n:.6f}")
