"""Entry point with the reference's `evaluate.py` arguments (evaluate.py:5-73): loads the pre-trained SwAV
head from `--out_dir` (`projection.pt`, written by pretrain.py), generates `--num_test_samples` images of
`--model`, computes their per-pixel codes and label maps on the B200 path (predict_swav_codes,
hfc_with_swav/swav_clustering.py:659-693) and - when a segmentor checkpoint is given - the one-shot
segmentor's label maps (src/one_shot_pipeline.py:655-666).  Writes `labels.pt` into `--out_dir/tests`.
The dataset side of the reference's pipeline (one-shot labels, IoU metrics, plots) is outside this path."""
import argparse
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse(argv=None):
    from ganecdotes_b200 import configs
    p = argparse.ArgumentParser(description="Label maps with a pre-trained clustering head")
    p.add_argument("--model", default='ffhq-256', choices=sorted(configs.MODELS), type=str)
    p.add_argument("--method", default='hfc_with_swav', choices=['hfc_with_swav', 'hfc_with_simclr', 'hfc_kmeans'],
                   type=str)
    p.add_argument("--out_dir", default="results/pretrain_default_ffhq/")
    p.add_argument("--expt_desc", default="Testing Clustering Model")
    p.add_argument("--num_test_samples", default=10, type=int)
    p.add_argument("--checkpoint", default=None, help="generator checkpoint (rosinality g_ema / BagGAN state dict)")
    p.add_argument("--segmentor", default=None, help="state dict of a fine-tuned OneShotSegmentor (optional)")
    p.add_argument("--n_class", default=12, type=int)
    p.add_argument("--seed", default=42, type=int)
    return p.parse_args(argv)


def main(argv=None):
    args = parse(argv)
    import numpy as np
    import torch
    from ganecdotes_b200 import configs
    from ganecdotes_b200.hfc_with_swav import OneShotSegmentor, SwAVClustering, engine as E
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    log = logging.getLogger("evaluate")
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    gen = configs.build_generator(args.model, args.checkpoint, 'cuda', args.seed)
    if args.method in ('hfc_with_simclr', 'hfc_kmeans'):
        z = torch.randn(args.num_test_samples, 512)
        t0 = time.time()
        with torch.no_grad():
            w = gen.style(z.cuda())
            if args.method == 'hfc_with_simclr':
                from ganecdotes_b200.hfc_with_simclr import SimCLRClustering
                obj = SimCLRClustering(model=gen, model_config=configs.model_config(args.model), logger=log, train=False,
                                       out_dir=args.out_dir, device='cuda', tb=None, **configs.simclr_config(args.model))
                if not hasattr(obj, 'projection'):
                    raise SystemExit(f"no projection.pt in {args.out_dir}: run pretrain.py --method hfc_with_simclr first")
                obj.projection.eval()
                _, labels = obj.predict_simclr_codes(w)
                out = {"latents": z, "code_labels": labels.cpu()}
            else:
                from ganecdotes_b200.hfc_kmeans import HFCPreprocessor
                obj = HFCPreprocessor(model=gen, model_config=configs.model_config(args.model), out_dir=args.out_dir,
                                      logger=log, train=False, **configs.kmeans_config(args.model))
                layer_labels = None
                for i in range(args.num_test_samples):        # the reference predicts one latent at a time (:628-631)
                    _, labs = obj.predict_hfc_vectors(w[i:i + 1])
                    layer_labels = [l.cpu() for l in labs] if layer_labels is None else \
                        [torch.cat([a, l.cpu()]) for a, l in zip(layer_labels, labs)]
                out = {"latents": z, "layer_labels": layer_labels, "code_labels": layer_labels[-1][:, 0].long()}
        torch.cuda.synchronize()
        test_dir = os.path.join(args.out_dir, 'tests')
        os.makedirs(test_dir, exist_ok=True)
        torch.save(out, os.path.join(test_dir, 'labels.pt'))
        log.info(f"{args.num_test_samples} label maps ({args.method}) in {time.time() - t0:.1f} s -> {test_dir}")
        return out
    cfg = configs.swav_config(args.model, args.method)
    swav = SwAVClustering(model=gen, model_config=configs.model_config(args.model), logger=log, train=False,
                          out_dir=args.out_dir, device='cuda', tb=None, **cfg)
    if not hasattr(swav, 'projection'):
        raise SystemExit(f"no projection.pt in {args.out_dir}: run pretrain.py first")
    z = torch.randn(args.num_test_samples, 512)
    t0 = time.time()
    with torch.no_grad():
        w = gen.style(z.cuda())
        preds, labels, planes = E.predict_codes(gen, swav.projection[0].weight.data, w, swav.mean_latent,
                                                swav.truncation, cfg['swav_args']['hlen'], want_planes=True)
        out = {"latents": z, "code_labels": labels.cpu()}
        if args.segmentor is not None:
            head = OneShotSegmentor(n_class=args.n_class, **configs.seg_args(args.model, args.method)).cuda().eval()
            head.load_state_dict(torch.load(args.segmentor, map_location='cuda'))
            out["labels"] = head.predict_labels(preds, planes).cpu()
    torch.cuda.synchronize()
    test_dir = os.path.join(args.out_dir, 'tests')
    os.makedirs(test_dir, exist_ok=True)
    torch.save(out, os.path.join(test_dir, 'labels.pt'))
    px = args.num_test_samples * labels.shape[1] * labels.shape[2]
    log.info(f"{args.num_test_samples} label maps ({px / (time.time() - t0) / 1e6:.1f} M pixels/s incl. host) -> {test_dir}")
    return out


if __name__ == "__main__":
    main()
