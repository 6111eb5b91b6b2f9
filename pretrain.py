"""Entry point with the reference's `pretrain.py` arguments (pretrain.py:5-74): pre-trains the SwAV
projection / prototype head of `--model` on the B200 path and writes `projection.pt` / `prototypes.pt`
into `--out_dir`, like the reference's `OneShotPipeline.run_trainer` does for the hfc_with_swav methods
(src/one_shot_pipeline.py:500-514).  Extras: `--checkpoint` (generator weights; default: seeded random
init, there are no checkpoints offline), `--num_epochs`, `--batch_latents`.  Under `torchrun` the latents
of a step are sharded over the ranks."""
import argparse
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse(argv=None):
    from ganecdotes_b200 import configs
    p = argparse.ArgumentParser(description="Pre-train the self-supervised clustering head for one-shot segmentation")
    p.add_argument("--model", default='ffhq-256', choices=sorted(configs.MODELS), type=str)
    p.add_argument("--method", default='hfc_with_swav', choices=['hfc_with_swav', 'hfc_with_simclr', 'hfc_kmeans'],
                   type=str, help="hfc_with_swav (the path) or one of the two baselines, like the reference's CLI")
    p.add_argument("--out_dir", default="results/pretrain_default_ffhq/")
    p.add_argument("--expt_desc", default="Testing Clustering Model")
    p.add_argument("--num_test_samples", default=10, type=int)
    p.add_argument("--checkpoint", default=None, help="generator checkpoint (rosinality g_ema / BagGAN state dict)")
    p.add_argument("--num_epochs", default=None, type=int, help="override swav_args['num_epochs'] (100)")
    p.add_argument("--batch_latents", default=1, type=int, help="latents per optimiser step (1 = the reference)")
    p.add_argument("--seed", default=42, type=int)
    return p.parse_args(argv)


def main(argv=None):
    args = parse(argv)
    import numpy as np
    import torch
    from ganecdotes_b200 import configs
    from ganecdotes_b200.hfc_with_swav import SwAVClustering
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    log = logging.getLogger("pretrain")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    torch.manual_seed(args.seed)          # the reference's seed_everything(42), lib/util/util.py:21-24
    np.random.seed(args.seed)
    os.makedirs(args.out_dir, exist_ok=True)
    log.info(f"{args.expt_desc}: model {args.model}, method {configs.method_for(args.model, args.method)}")
    gen = configs.build_generator(args.model, args.checkpoint, 'cuda', args.seed)
    if args.method == 'hfc_with_simclr':        # baseline/hfc_with_simclr (src/one_shot_pipeline.py:207-209)
        from ganecdotes_b200.hfc_with_simclr import SimCLRClustering
        cfg = configs.simclr_config(args.model)
        if args.num_epochs is not None:
            cfg['simclr_args']['num_iters'] = args.num_epochs
        obj = SimCLRClustering(model=gen, model_config=configs.model_config(args.model), logger=log, train=True,
                               out_dir=args.out_dir, device='cuda', tb=None, **cfg)
        t0 = time.time()
        obj.pretrain(None, num_test_samples=args.num_test_samples)
        torch.cuda.synchronize()
        log.info(f"SimCLR pre-training done in {time.time() - t0:.1f} s -> {args.out_dir}")
        return obj
    if args.method == 'hfc_kmeans':             # baseline/hfc_kmeans (src/one_shot_pipeline.py:211-219, 484-488)
        from ganecdotes_b200.hfc_kmeans import HFCPreprocessor
        obj = HFCPreprocessor(model=gen, model_config=configs.model_config(args.model), out_dir=args.out_dir, logger=log,
                              train=True, **configs.kmeans_config(args.model))
        t0 = time.time()
        with torch.no_grad():
            one_shot_latent = gen.style(torch.randn(1, 512).cuda())      # no dataset offline: a seeded random latent
        obj.train_hfc_model(one_shot_latent)
        torch.cuda.synchronize()
        log.info(f"k-means models fitted in {time.time() - t0:.1f} s -> {args.out_dir}")
        return obj
    cfg = configs.swav_config(args.model, args.method)
    if args.num_epochs is not None:
        cfg['swav_args']['num_epochs'] = args.num_epochs
    cfg['swav_args']['batch_latents'] = args.batch_latents
    swav = SwAVClustering(model=gen, model_config=configs.model_config(args.model), logger=log, train=True,
                          out_dir=args.out_dir, device='cuda', tb=None, **cfg)
    t0 = time.time()
    swav.pretrain(None, num_test_samples=args.num_test_samples)
    torch.cuda.synchronize()
    log.info(f"pre-training done in {time.time() - t0:.1f} s -> {args.out_dir}")
    return swav


if __name__ == "__main__":
    main()
