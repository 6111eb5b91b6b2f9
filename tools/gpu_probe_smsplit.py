"""How do the two kernel families scale with the number of SMs they get?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
n, k = 160000, 5000
s = (0.05 * torch.randn(n, k, device="cuda"))
ws = L.SinkhornWorkspace(k, "cuda")
a = torch.randn(n, 5376, device="cuda").to(torch.bfloat16)
w = torch.randn(512, 5376, device="cuda").to(torch.bfloat16)
out = torch.empty(n, 512, device="cuda")

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for ctas in (148, 116, 100, 84, 64, 48, 32):
    L.set_sm_budget(ctas, ctas)
    u = L.sinkhorn_pass(s, 200.0, True, None, None, None, n, ws).clone()
    t_sk = timeit(lambda: L.sinkhorn_pass(s, 200.0, False, u, None, None, n, ws))
    t_g3 = timeit(lambda: L.gemm(a, a, w, w, n, 512, 5376, 3, out=out))
    t_g1 = timeit(lambda: L.gemm(a, None, w, None, n, 512, 5376, 1, out=out))
    print(f"ctas={ctas:4d}  sinkhorn_pass {t_sk:7.3f} ms {n*k*4/t_sk/1e6:8.1f} GB/s   gemm3 {t_g3:7.3f} ms {2*n*512*5376/t_g3/1e9:7.1f} TF   gemm1 {t_g1:7.3f} ms {2*n*512*5376/t_g1/1e9:7.1f} TF", flush=True)

# concurrent: gemm on 100 SMs + sinkhorn on 48 SMs on two streams
for g_ctas, s_ctas in ((100, 48), (110, 38), (116, 32), (90, 58)):
    st_t, st_h = torch.cuda.Stream(), torch.cuda.Stream()
    L.set_sm_budget(g_ctas, s_ctas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st_t.wait_stream(torch.cuda.current_stream()); st_h.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st_t):
        for _ in range(4): L.gemm(a, a, w, w, n, 512, 5376, 3, out=out)
    with torch.cuda.stream(st_h):
        for _ in range(16): L.sinkhorn_pass(s, 200.0, False, u, None, None, n, ws)
    torch.cuda.current_stream().wait_stream(st_t); torch.cuda.current_stream().wait_stream(st_h)
    e1.record(); torch.cuda.synchronize()
    print(f"concurrent gemm3 x4 on {g_ctas} + sinkhorn x16 on {s_ctas}: {e0.elapsed_time(e1):.3f} ms", flush=True)
L.set_sm_budget(0, 0)
