"""Probe of the pixel x prototype score GEMM (n=160000, c=512, k=5000): time per launch for
the bf16x3 / bf16x1 / fp16x1 operand modes, with and without the fused exp column sums, and
with the epilogue stages switched off (GX_UMMA_DEBUG bits) to see what bounds the tile."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from ganecdotes_b200 import _lib as L
    torch.manual_seed(0)
    n, c, k = 160000, 512, 5000
    z = torch.randn(n, c, device="cuda") / 22
    w = torch.randn(k, c, device="cuda") / 22
    zb, wb = z.to(torch.bfloat16), w.to(torch.bfloat16)
    zh, wh = z.half(), w.half()
    b = torch.randn(k, device="cuda") * 0.01
    out = torch.empty(n, k, device="cuda")
    u = torch.zeros(k, device="cuda")

    def run(name, fn, flops=2.0 * n * k * c):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{os.environ.get('GX_UMMA_DEBUG', '0')} {name:28s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  "
              f"store {n * k * 4 / ms / 1e6:7.1f} GB/s")

    run("bf16x3", lambda: L.gemm(zb, zb, wb, wb, n, k, c, 3, out=out, bias=b))
    run("bf16x3 + colexp", lambda: L.gemm(zb, zb, wb, wb, n, k, c, 3, out=out, bias=b, colexp=(u, 20.0)))
    run("bf16x1", lambda: L.gemm(zb, None, wb, None, n, k, c, 1, out=out, bias=b))
    run("bf16x1 m128", lambda: L.gemm(zb, None, wb, None, n, k, c, 1, out=out, bias=b, force_m128=True))
    run("fp16x1 + colexp", lambda: L.gemm(zh, None, wh, None, n, k, c, 1, out=out, bias=b, colexp=(u, 20.0)))
    run("bf16x3 pair + colexp",
        lambda: L.gemm(zb, zb, wb, wb, n, k, c, 3, out=out, bias=b, colexp=(u, 20.0), pair=True))
    run("bf16x1 pair", lambda: L.gemm(zb, None, wb, None, n, k, c, 1, out=out, bias=b, pair=True))
    run("fp16x1 pair + colexp",
        lambda: L.gemm(zh, None, wh, None, n, k, c, 1, out=out, bias=b, colexp=(u, 20.0), pair=True))
    run("fp16x1 m128 + colexp",
        lambda: L.gemm(zh, None, wh, None, n, k, c, 1, out=out, bias=b, colexp=(u, 20.0), force_m128=True))
else:
    for dbg in os.environ.get("GX_PROBE_MODES", "0,1,2").split(","):
        env = dict(os.environ, GX_UMMA_DEBUG=dbg)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=True)
