"""Multi-rank parity check (run under torchrun, 1 rank per GPU): a step over a global batch sharded across ranks
(distributed Sinkhorn over the NVLink exchange + gradient all-reduce) must equal the single-process step over the
same global batch, and the exchange over peer memory must give the NCCL all-reduce's result.  Two geometries: the
16^2 generator (K = 48) and a K = 5000 head (the full-width Sinkhorn kernels)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganecdotes_b200 import _lib as L  # noqa: E402
from ganecdotes_b200.hfc_with_swav import engine as E  # noqa: E402
from ganecdotes_b200.hfc_with_swav.swav_clustering import SwAVClustering  # noqa: E402
from ganecdotes_b200.stylegan2.model import Generator  # noqa: E402


def one_case(gen, mean_latent, dev, rank, world, hlen, c, k, patch, npatch, eps, temp, seed):
    torch.manual_seed(seed)
    wp = torch.randn(c, hlen) / hlen ** 0.5
    wk = torch.randn(k, c)
    bk = 0.05 * torch.randn(k)
    cfg = E.StepConfig(hlen=hlen, patch_size=patch, num_patches=npatch, niters=10, eps=eps, temperature=temp,
                       truncation=0.7, perturb_std=[1.0, 0.5, 1.0])
    b = 2 * world
    g = torch.Generator().manual_seed(seed + 1)
    rs = np.random.RandomState(seed + 1)
    view = lambda: E.ViewDraws(layer_no=[int(rs.randint(3)) for _ in range(b)],
                               pert_z=torch.randn(b, 6, 64, generator=g),
                               angle=[float(rs.uniform(-10, 10)) for _ in range(b)],
                               flip=[bool(rs.rand() < 0.5) for _ in range(b)])
    draws = E.StepDraws(z=torch.randn(b, 64, generator=g), view_s=view(), view_t=view(),
                        perms=[[torch.randperm(256, generator=g) for _ in range(b)] for _ in range(npatch)])
    mk = lambda: E.SwavHead(wp.clone().to(dev), wk.clone().to(dev), bk.clone().to(dev), 0.01, 0.9, 0.01, 3, 3)
    shard = SwAVClustering.shard_draws(draws, rank, world)
    res = {}
    for mode in ("ll", "nccl"):
        group = E.DistGroup(dist.group.WORLD, rank, world)
        if mode == "ll":
            group.ensure_ll(k, dev)
        os.environ["GX_SINKHORN_EXCHANGE"] = mode
        head = mk()
        losses = [E.swav_train_step(gen, head, mean_latent, shard, cfg, group) for _ in range(2)]   # 2 steps: parity of
        torch.cuda.synchronize()                                                                  # the blocks wraps
        if group.ll is not None:
            group.ll.check()
            dist.barrier()
            group.ll.close()
        res[mode] = (head, [l.item() for l in losses])
    ref = mk()
    ref_losses = [E.swav_train_step(gen, ref, mean_latent, draws, cfg, None).item() for _ in range(2)]
    torch.cuda.synchronize()
    ok = True
    why = []

    def check(name, cond, val):
        nonlocal ok
        if not cond:
            ok = False
            why.append(f"{name}={val:.3e}")

    for mode, (head, losses) in res.items():
        for i, (l, lr) in enumerate(zip(losses, ref_losses)):
            check(f"{mode}.loss{i}", abs(l - lr) < 1e-4 * abs(lr), abs(l - lr) / abs(lr))
        for nm, a, r in zip(("w_proj", "w_proto", "b_proto"), (head.w_proj, head.w_proto, head.b_proto),
                            (ref.w_proj, ref.w_proto, ref.b_proto)):
            err = ((a - r).abs() / (1e-6 + 1e-4 * r.abs())).max().item()
            check(f"{mode}.{nm}", err <= 1.0, err)
        for nm, a, r in zip(("g_proj", "g_proto", "g_bias"), (head.g_proj, head.g_proto, head.g_bias),
                            (ref.g_proj, ref.g_proto, ref.g_bias)):
            rel = ((a - r).norm() / r.norm()).item()
            check(f"{mode}.{nm}", rel < 1e-3, rel)
    # the two transports reduce the same numbers: weights agree to the rounding of the summation order
    for nm, a, r in zip(("w_proj", "w_proto"), (res["ll"][0].w_proj, res["ll"][0].w_proto),
                        (res["nccl"][0].w_proj, res["nccl"][0].w_proto)):
        err = ((a - r).abs() / (1e-6 + 1e-4 * r.abs())).max().item()
        check(f"ll_vs_nccl.{nm}", err <= 1.0, err)
    # replicas stay bit-identical: every rank holds the same weights after the LL steps
    wsum = res["ll"][0].w_proto.double().sum().reshape(1)
    gathered = [torch.zeros_like(wsum) for _ in range(world)]
    dist.all_gather(gathered, wsum)
    check("replicas_identical", all(torch.equal(gathered[0], t) for t in gathered),
          max((t - gathered[0]).abs().item() for t in gathered))
    if why:
        print(f"[rank {rank}] K={k} failed checks: {' '.join(why)}", flush=True)
    return ok, res["ll"][1], ref_losses


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    L.load()
    torch.manual_seed(3)
    gen = Generator(16, 64, 2).to(dev)
    mean_latent = gen.style(torch.randn(64, 64).to(dev)).mean(0, keepdim=True)
    all_ok = True
    for (c, k, patch, npatch, eps, temp) in ((64, 48, 80, 2, 0.02, 0.02), (64, 5000, 200, 3, 0.05, 0.05)):
        ok, losses, ref_losses = one_case(gen, mean_latent, dev, rank, world, 2560, c, k, patch, npatch, eps, temp, 5 + k)
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = all_ok and flag.item() == 1.0
        if rank == 0:
            print(f"DIST_CHECK world={world} K={k} losses={losses} ref={ref_losses} "
                  f"{'OK' if flag.item() == 1.0 else 'MISMATCH'}", flush=True)
    # the public class under torchrun: SwAVClustering.pretrain shards `batch_latents` over the ranks; deliberately
    # DIFFERENT seeds per rank (the class broadcasts weights / mean latent and re-seeds the draw generators)
    import types
    torch.manual_seed(100 + rank)
    np.random.seed(100 + rank)
    mc = types.SimpleNamespace(num_latents_for_mean=64, truncation=0.7, latent_dim=64, image_size=16)
    cfg = dict(perturb_args=dict(truncation=0.7, n_layers=3, n_samples=1, layer_no=None, perturb_std=[1.0, 0.5, 1.0]),
               swav_args=dict(num_epochs=3, num_samples=1, num_patches=2, sampling_method='random', patch_size=100,
                              hf_interp='nearest', warmup_epochs=3, start_warmup=0.01, use_scheduler=True, base_lr=0.01,
                              final_lr=0.0001, trust_coeff=0.01, freeze_prototype_niters=313,
                              train_args=dict(lr=0.01, momentum=0.9), projn_nw='linear', temperature=0.02, nprototypes=48,
                              nclasses=64, hlen=2560, add_local_loss=False, plot_test_images=False, epoch_print_freq=1,
                              max_masks=4, batch_latents=2 * world),
               sinkhorn_args=dict(source_pdf='uniform', niters=10, eps=0.02), train=True, layer_hf_dim=[512, 1024, 1024])
    losses = []
    tb = types.SimpleNamespace(add_scalar=lambda name, val, step: losses.append(float(val)))
    obj = SwAVClustering(gen, mc, out_dir=None, device=str(dev), tb=tb, **cfg)
    obj.pretrain(None, num_test_samples=1)
    torch.cuda.synchronize()
    sig = torch.stack([obj.projection[0].weight.double().sum(), obj.prototype.weight.double().sum(),
                       torch.tensor(losses[-1], dtype=torch.float64, device=dev)])
    gathered = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(gathered, sig)
    ok_cls = all(torch.equal(gathered[0], t) for t in gathered) and all(np.isfinite(losses)) and len(losses) == 3
    all_ok = all_ok and ok_cls
    if rank == 0:
        print(f"DIST_CHECK world={world} SwAVClustering.pretrain (batch_latents={2 * world}, ranks seeded differently) "
              f"losses={losses} replicas_identical={ok_cls} {'OK' if ok_cls else 'MISMATCH'}", flush=True)
        print("DIST_CHECK", "OK" if all_ok else "MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
