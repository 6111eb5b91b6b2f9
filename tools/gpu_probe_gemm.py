"""First-contact probe of the tcgen05 engine on a B200: runs each GEMM / conv variant and
prints max errors against fp64 torch.  Run under `timeout`; a protocol bug traps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"


def planes(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def run_gemm(m, n, k, passes, a_mn, b_mn, split_k=1, bias=False, block_n=0, stages=0):
    a = torch.randn(m, k, device=dev)
    b = torch.randn(n, k, device=dev)
    ah, al = planes(a.t().contiguous() if a_mn else a)
    bh, bl = planes(b.t().contiguous() if b_mn else b)
    bv = torch.randn(n, device=dev) if bias else None
    out = L.gemm(ah, al if passes == 3 else None, bh, bl if passes == 3 else None, m, n, k, passes, bias=bv,
                 a_mn=a_mn, b_mn=b_mn, split_k=split_k, block_n=block_n, stages=stages)
    torch.cuda.synchronize()
    if passes == 3:
        ar = (ah.double() + al.double()); br = (bh.double() + bl.double())
    else:
        ar = ah.double(); br = bh.double()
    if a_mn: ar = ar.t()
    if b_mn: br = br.t()
    ref = ar @ br.t()
    if bias: ref = ref + bv.double()
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    ref32 = a.double() @ b.double().t() + (bv.double() if bias else 0)
    err32 = (out.double() - ref32).abs().max().item()
    print(f"gemm m={m} n={n} k={k} passes={passes} a_mn={int(a_mn)} b_mn={int(b_mn)} split={split_k} bias={int(bias)} "
          f"bn={block_n}: err_vs_planes={err:.3e} err_vs_fp32={err32:.3e} scale={scale:.2f}", flush=True)
    return err / scale


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    if which in ("all", "kk"):
        run_gemm(128, 128, 64, 1, False, False, block_n=128)
        run_gemm(128, 128, 256, 1, False, False, block_n=128)
        run_gemm(256, 256, 512, 1, False, False)
        run_gemm(256, 256, 512, 3, False, False)
        run_gemm(1000, 520, 5376, 3, False, False, bias=True)
        run_gemm(777, 5000, 512, 3, False, False, bias=True)
        run_gemm(300, 72, 200, 1, False, False)
    if which in ("all", "mn"):
        run_gemm(128, 128, 64, 1, True, False, block_n=128)
        run_gemm(128, 128, 64, 1, False, True, block_n=128)
        run_gemm(256, 256, 512, 1, True, True)
        run_gemm(520, 512, 4096, 3, True, True, split_k=4)
        run_gemm(5000, 512, 3000, 1, True, True, split_k=7)
    print("probe done", flush=True)
