"""Probe of the Sinkhorn passes through the 16-bit cache (gx_sinkhorn_pass_cached): accuracy of log a and of the
codes against the fp32 passes and the fp64 oracle, and the per-pass times at the full ffhq-256 size.
usage: python tools/gpu_probe_sk16.py [n k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
from ganecdotes_b200.hfc_with_swav import engine as E
from oracle import ganecdotes_oracle as O


def accuracy(n, k, std, eps=0.005, r=None, c=None):
    torch.manual_seed(n + k)
    s = (std * torch.randn(n, k)).cuda()
    ws = L.SinkhornWorkspace(k, "cuda")
    la32 = E.sinkhorn_log_a(s, 10, eps, ws, n)
    la16 = E.sinkhorn_log_a(s, 10, eps, ws, n, cache16=True)
    q32, q16 = L.sinkhorn_q(s, 1 / eps, la32), L.sinkhorn_q(s, 1 / eps, la16)
    ref = O.sinkhorn_knopp(s.cpu().double(), 10, eps)
    m = ref > 1e-9

    def rel(q):
        return ((q.cpu().double() - ref).abs() / ref)[m].max().item()
    print(f"n={n} k={k} std={std}: |log a16 - log a32| max {float((la16 - la32).abs().max()):.2e}; "
          f"codes vs fp64 oracle (max rel over q > 1e-9): fp32 passes {rel(q32):.2e}, cached passes {rel(q16):.2e}")


def timing(n, k, eps=0.005):
    torch.manual_seed(0)
    s = (0.05 * torch.randn(n, k)).cuda()
    ws = L.SinkhornWorkspace(k, "cuda")
    cache = ws.cache16(0, n)
    u0 = L.sinkhorn_pass(s, 1 / eps, True, None, None, None, n, ws).clone()

    def t(fn, reps=10):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    rev = [0]

    def p32():
        rev[0] ^= 1
        L.sinkhorn_pass_parts(s, 1 / eps, False, u0, None, None, n, ws, reverse=rev[0])

    def pw():
        rev[0] ^= 1
        L.sinkhorn_pass_cached_parts(s, 1 / eps, u0, None, None, n, ws, cache, True, reverse=rev[0])

    def pr():
        rev[0] ^= 1
        L.sinkhorn_pass_cached_parts(s, 1 / eps, u0, None, None, n, ws, cache, False, reverse=rev[0])
    t32, tw, tr = t(p32), t(pw), t(pr)
    gb = n * k / 1e9
    print(f"n={n} k={k}: fp32 pass {t32:.4f} ms ({4 * gb / t32 * 1e3:.0f} GB/s); pass + cache write {tw:.4f} ms "
          f"({6 * gb / tw * 1e3:.0f} GB/s); cached pass {tr:.4f} ms ({2 * gb / tr * 1e3:.0f} GB/s)")

    def chain(c16):
        E.sinkhorn_log_a(s, 10, eps, ws, n, u_first=u0, cache16=c16)
    print(f"   10-iteration chain (first marginals given): fp32 {t(lambda: chain(False), 4):.3f} ms, "
          f"cached {t(lambda: chain(True), 4):.3f} ms")


if __name__ == "__main__":
    if len(sys.argv) > 2:
        timing(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for n, k, std in [(1000, 5000, 0.05), (4096, 1024, 0.05), (8192, 2048, 0.05), (20000, 5000, 0.05),
                          (3001, 4000, 0.02), (777, 520, 0.05)]:
            accuracy(n, k, std)
        timing(160000, 5000)
        timing(40000, 4000, 0.01)
