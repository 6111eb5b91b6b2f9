"""Probe of the streaming (TMA ring) kernels at sizes where the ring wraps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
from ganecdotes_b200.hfc_with_swav import engine as E

torch.manual_seed(0)
n, k = int(sys.argv[1]), int(sys.argv[2])
s = (0.05 * torch.randn(n, k)).cuda()
t = (0.05 * torch.randn(n, k)).cuda()
ws = L.SinkhornWorkspace(k, "cuda")
print("pass first", flush=True)
u = L.sinkhorn_pass(s, 200.0, True, None, None, None, n, ws); torch.cuda.synchronize()
print("pass 2", flush=True)
u = L.sinkhorn_pass(s, 200.0, False, u, None, None, n, ws); torch.cuda.synchronize()
la_s = E.sinkhorn_log_a(s, 10, 0.005, ws, n); torch.cuda.synchronize()
la_t = E.sinkhorn_log_a(t, 10, 0.005, ws, n); torch.cuda.synchronize()
print("loss", flush=True)
out = L.swav_loss(s, t, 200.0, 100.0, la_s, la_t, 1.0 / n)
torch.cuda.synchronize()
print("loss value", out[0].sum().item() / n, flush=True)
q = L.sinkhorn_q(s, 200.0, la_s)
print("q row sums", q.sum(1).min().item(), q.sum(1).max().item(), "col sums*K", (q.sum(0) * k / n).min().item(), (q.sum(0) * k / n).max().item())
