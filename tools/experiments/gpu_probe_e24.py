"""Timing of the E24 path at the bench shapes (N = 160000 rows, K = 5000, C = 512, eps = 0.005):
score GEMM with / without the E24 planes (and without the fp32 store), Sinkhorn pass on S vs on the planes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
from ganecdotes_b200.hfc_with_swav import engine as E

torch.manual_seed(0)
n, k, c, eps = int(sys.argv[1]) if len(sys.argv) > 1 else 160000, 5000, 512, 0.005
PEAK = 6529.1


def med_ms(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


zn = torch.nn.functional.normalize(torch.randn(n, c, device="cuda"), dim=1)
wk = torch.nn.functional.normalize(torch.randn(k, c, device="cuda"), dim=1)
bias = 0.02 * torch.randn(k, device="cuda")
zh, zl = L.split_planes(zn)
wh, wl = L.split_planes(wk)
s = torch.empty(n, k, device="cuda")
u = torch.zeros(k, device="cuda")
e24 = L.e24_planes(n, k, "cuda")
sc = 1.4426950408889634 / eps


def gemm(**kw):
    u.zero_()
    return L.gemm(zh, zl, wh, wl, n, k, c, 3, out=s, bias=bias, colexp=(u, sc), pair=True, **kw)


for name, kw in (("S only", {}), ("S + E24", dict(e24=e24)), ("E24 only", dict(e24=e24, e24_only=True))):
    ms = med_ms(lambda: gemm(**kw))
    print(f"score GEMM bf16x3, {name:9s}: {ms:.3f} ms  ({2.0 * n * k * c / ms / 1e9:.0f} TFLOP/s algorithmic)")
gemm(e24=e24)
torch.cuda.synchronize()
ws = L.SinkhornWorkspace(k, "cuda")
u0 = u.clone()
ms = med_ms(lambda: L.sinkhorn_pass(s, 1.0 / eps, False, u0, None, None, n, ws))
print(f"sinkhorn_pass     (fp32 S, 4 B/score): {ms:.4f} ms  {4.0 * n * k / ms / 1e6:.0f} GB/s ({4.0 * n * k / ms / 1e6 / PEAK:.0%})")
ua = L.sinkhorn_pass(s, 1.0 / eps, False, u0, None, None, n, ws).clone()
ms = med_ms(lambda: L.sinkhorn_pass_e24(e24, False, u0, None, None, n, ws))
print(f"sinkhorn_pass_e24 (planes, 3 B/score): {ms:.4f} ms  {3.0 * n * k / ms / 1e6:.0f} GB/s ({3.0 * n * k / ms / 1e6 / PEAK:.0%})")
ub = L.sinkhorn_pass_e24(e24, False, u0, None, None, n, ws).clone()
print("second-pass u: max rel diff e24 vs fp32:", ((ua - ub).abs() / ua).max().item())
la_a = E.sinkhorn_log_a(s, 10, eps, ws, n, u_first=u0).clone()
la_b = E.sinkhorn_log_a(s, 10, eps, ws, n, u_first=u0, e24=e24).clone()
print("log a after 10 iterations: max abs diff:", (la_a - la_b).abs().max().item())
qa = torch.softmax(s[:2048].double() / eps + la_a.double(), dim=1)
qb = torch.softmax(s[:2048].double() / eps + la_b.double(), dim=1)
m = qa > 1e-4
print("codes (entries > 1e-4): rel diff rms / max:", (((qa - qb) / qa)[m] ** 2).mean().sqrt().item(),
      ((qa - qb).abs() / qa)[m].max().item())
