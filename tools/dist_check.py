"""Multi-rank parity check (run under torchrun, 1 rank per GPU): a step over a global
batch sharded across ranks (distributed Sinkhorn + gradient all-reduce) must equal the
single-process step over the same global batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganecdotes_b200.hfc_with_swav import engine as E  # noqa: E402
from ganecdotes_b200.hfc_with_swav.swav_clustering import SwAVClustering  # noqa: E402
from ganecdotes_b200.stylegan2.model import Generator  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(3)
    gen = Generator(16, 64, 2).to(dev)
    mean_latent = gen.style(torch.randn(64, 64).to(dev)).mean(0, keepdim=True)
    hlen, c, k, patch, npatch = 2560, 64, 48, 80, 2
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    cfg = E.StepConfig(hlen=hlen, patch_size=patch, num_patches=npatch, niters=10, eps=0.02, temperature=0.02,
                       truncation=0.7, perturb_std=[1.0, 0.5, 1.0])
    b = 2 * world
    g = torch.Generator().manual_seed(5)
    rs = np.random.RandomState(5)
    view = lambda: E.ViewDraws(layer_no=[int(rs.randint(3)) for _ in range(b)],
                               pert_z=torch.randn(b, 6, 64, generator=g),
                               angle=[float(rs.uniform(-10, 10)) for _ in range(b)],
                               flip=[bool(rs.rand() < 0.5) for _ in range(b)])
    draws = E.StepDraws(z=torch.randn(b, 64, generator=g), view_s=view(), view_t=view(),
                        perms=[[torch.randperm(256, generator=g) for _ in range(b)] for _ in range(npatch)])
    group = E.DistGroup(dist.group.WORLD, rank, world)
    head = E.SwavHead(wp.clone().to(dev), wk.clone().to(dev), bk.clone().to(dev), 0.01, 0.9, 0.01, 3, 3)
    loss = E.swav_train_step(gen, head, mean_latent, SwAVClustering.shard_draws(draws, rank, world), cfg, group)
    ref = E.SwavHead(wp.clone().to(dev), wk.clone().to(dev), bk.clone().to(dev), 0.01, 0.9, 0.01, 3, 3)
    loss_ref = E.swav_train_step(gen, ref, mean_latent, draws, cfg, None)
    torch.cuda.synchronize()
    ok = abs(loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    for a, r in zip((head.w_proj, head.w_proto, head.b_proto), (ref.w_proj, ref.w_proto, ref.b_proto)):
        ok = ok and torch.allclose(a, r, rtol=1e-4, atol=1e-6)
    for a, r in zip((head.g_proj, head.g_proto, head.g_bias), (ref.g_proj, ref.g_proto, ref.g_bias)):
        ok = ok and ((a - r).norm() / r.norm()).item() < 1e-3
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DIST_CHECK world={world} loss={loss.item():.6f} ref={loss_ref.item():.6f} "
              f"{'OK' if flag.item() == 1.0 else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
