"""A/B timing of the fused loss kernel at the ffhq-256 shape (160000 x 5000 scores per view): the lean instantiation
(single bf16 dS planes) against the general one (GX_LOSS_NOFAST=1), alternating in one process; median of 15 launches
per round, CUDA events around each launch."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
n, k = 160000, int(os.environ.get("K", 5000))
inv_eps = float(os.environ.get("INV_EPS", 200.0))
s = 0.05 * torch.randn(n, k, device="cuda")
t = 0.05 * torch.randn(n, k, device="cuda")
la = torch.zeros(k, device="cuda")


def run():
    return L.swav_loss(s, t, inv_eps, 100.0, la, la, 1.0 / n)


def med(nrep=15):
    ts = []
    for _ in range(nrep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out


for _ in range(5):
    run()
for rnd in range(3):
    for mode in ("fast", "general"):
        if mode == "general":
            os.environ["GX_LOSS_NOFAST"] = "1"
        else:
            os.environ.pop("GX_LOSS_NOFAST", None)
        ms, out = med()
        print(f"round {rnd} {mode:8s}: {ms:.3f} ms  {n * k * 12 / ms / 1e6:.0f} GB/s  loss={out[0].sum().item() / n:.6f}")
