"""Per-layer timing of the k-means assignment routes at the config-5 shapes (16 images): the fused kernel
(gx_kmeans_assign_mma) against the GEMM route and the per-centre SIMT kernel; median of 20 launches, CUDA events.
`python tools/gpu_probe_kmeans.py one` runs two launches of the two largest layers only (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
PEAK = 6529.1
B = 16
LAYERS = [(8, 1024, 4), (16, 1024, 8), (32, 1024, 16), (64, 1024, 32), (128, 512, 64)]


def med_ms(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


one = len(sys.argv) > 1 and sys.argv[1] == "one"
if len(sys.argv) > 1 and sys.argv[1] == "pattern":
    # same bytes (268 MB), different row lengths: is the row-fragmented access pattern what limits the fused kernels?
    for n, c1, c2, k in ((4194304, 64, 0, 32), (262144, 512, 512, 32), (524288, 256, 256, 64)):
        x1 = torch.randn(n, c1, device=dev)
        x2 = torch.randn(n, c2, device=dev) if c2 else None
        cen = torch.randn(k, c1 + c2, device=dev)
        nbytes = 4.0 * n * (c1 + c2)
        out = []
        for name, t in (("fused", True),):
            ms = med_ms(lambda: L.kmeans_assign(x1, cen, x2, tensor=t))
            out.append(f"{name} {ms * 1e3:8.1f} us ({nbytes / ms / 1e6 / PEAK:.2f})")
        print(f"n={n:8d} c1={c1:5d} c2={c2:5d} k={k}  " + "  ".join(out))
        del x1, x2
    sys.exit(0)
tot = {}
for h, c, k in (LAYERS[3:] if one else LAYERS):
    n = B * h * h
    x1 = torch.randn(n, c // 2, device=dev) + 0.5
    x2 = torch.randn(n, c // 2, device=dev) + 0.5
    cen = torch.randn(k, c, device=dev) + 0.5
    if one:
        for _ in range(2):
            L.kmeans_assign(x1, cen, x2, tensor=True)
        continue
    nbytes = 4.0 * n * c
    res = {}
    for name, t in (("fused", True), ("gemm", "gemm"), ("simt", False)):
        if name == "simt" and n * k > 3e6:
            continue
        ms = med_ms(lambda: L.kmeans_assign(x1, cen, x2, tensor=t))
        res[name] = ms
        tot[name] = tot.get(name, 0.0) + ms
    print(f"{h:4d}^2 C={c} K={k:3d} n={n:7d} {nbytes / 1e6:7.1f} MB  " +
          "  ".join(f"{nm} {ms * 1e3:8.1f} us ({nbytes / ms / 1e6 / PEAK:.2f})" for nm, ms in res.items()))
    a = L.kmeans_assign(x1, cen, x2, tensor=True)
    b = L.kmeans_assign(x1, cen, x2, tensor="gemm")
    print("      labels differ from the GEMM route at", int((a != b).sum()), "of", n)
if not one:
    print("sum over layers:", {k: round(v, 4) for k, v in tot.items()})
torch.cuda.synchronize()
print("done")
