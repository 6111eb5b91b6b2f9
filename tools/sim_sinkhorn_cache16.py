"""CPU (fp64) simulation of the Sinkhorn passes on the 16-bit cache: how far the codes
softmax_k(S/eps + log a) move from the exact ten iterations, as a function of the iteration that writes the cache and of
the shape of the score rows.  No GPU, no library: the scheme is restated here (the same arithmetic as
sinkhorn_pass_kernel<.., 1> / sinkhorn_pass16_kernel, with the fp16 rounding of the plane as the only inexact step).
usage: python tools/sim_sinkhorn_cache16.py [n k]      (default 20000 5000; ~6 minutes on 16 cores)"""
import sys

import torch


def codes_error(s, write_iter, eps=0.005, niters=10):
    e = torch.exp(s.double() / eps)
    n, k = e.shape
    r, c = 1.0 / k, 1.0 / n
    ux = e.sum(0)
    for _ in range(1, niters):                               # the exact iteration
        ux = (e * (c / (e @ (r / ux))).unsqueeze(1)).sum(0)
    u = e.sum(0)
    for _ in range(1, write_iter):                           # fp32-score passes before the cache exists
        u = (e * (c / (e @ (r / u))).unsqueeze(1)).sum(0)
    aw = r / u
    p = e * aw
    t = p.sum(1)
    e16 = (32768.0 * p / t.unsqueeze(1)).to(torch.float16).double()
    u = (p * (c / t).unsqueeze(1)).sum(0) / aw               # the writing pass itself uses its exact terms
    rho_max = 0.0
    for _ in range(write_iter + 1, niters):
        rho = (r / u) / aw
        rho_max = max(rho_max, rho.max().item())
        u = (e16 * (c / (e16 @ rho)).unsqueeze(1)).sum(0) / aw
    qx = torch.softmax(s.double() / eps + torch.log(r / ux), 1)
    qd = torch.softmax(s.double() / eps + torch.log(r / u), 1)
    big = qx > 1e-9
    return ((qd - qx).abs() / qx)[big].max().item(), rho_max, (e16 == 0).double().mean().item()


def main():
    n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20000, 5000)
    torch.manual_seed(0)
    rows = torch.arange(n)
    cases = {"random, sigma(S) = 0.044 (random init)": 0.044 * torch.randn(n, k)}
    for boost in (0.12, 0.15, 0.18, 0.25, 0.30):
        s = 0.02 * torch.randn(n, k)
        s[rows, torch.randint(0, k, (n,))] += boost
        cases[f"sharp: one prototype +{boost:.2f} ({boost / 0.005:.0f} nats) per row"] = s
    s = 0.02 * torch.randn(n, k)
    s[rows, torch.randint(0, 50, (n,))] += 0.2
    cases["sharp on 50 prototypes"] = s
    s = 0.03 * torch.randn(n, k)
    s[rows, torch.randint(0, 300, (n,))] += 0.25
    cases["clustered on 300 prototypes"] = s
    s = 0.03 * torch.randn(n, k)
    s[rows, (torch.rand(n) ** 3 * k).long()] += 0.2
    cases["skewed prototype usage"] = s
    cases["column bias"] = 0.044 * torch.randn(n, k) + 0.03 * torch.randn(1, k)
    cases["flat, sigma(S) = 0.01"] = 0.01 * torch.randn(n, k)
    print(f"# {n} x {k} scores, eps 0.005, 10 iterations: max relative error of the codes (entries > 1e-9) against the exact "
          f"iteration, by the iteration that writes the 16-bit cache; rho = a / a_write over the cached iterations")
    for name, s in cases.items():
        out = []
        for w in (1, 2, 3):
            err, rho, zero = codes_error(s, w)
            out.append(f"write@{w}: {err:.2e} (rho <= {rho:.0e}, {100 * zero:.0f} % of the plane is 0)")
        print(f"{name:52s} max S/eps {float((s / 0.005).max()):4.0f}   " + "   ".join(out), flush=True)


if __name__ == "__main__":
    main()
