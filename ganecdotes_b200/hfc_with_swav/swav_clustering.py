"""Drop-in for the reference's `hfc_with_swav/swav_clustering.py::SwAVClustering`
(same constructor, public methods, attributes and artefacts) running on the fused
sm_100a engine (`engine.py`).

Differences that are deliberate and documented:
  * the discarded first synthesis of every view (ref :603-607) and the ten mapping passes
    multiplied by sigma = 0 (ref image_augmentor.py:42-53) are skipped; the random draws
    they consume are still drawn, in the reference's order, so a seeded run follows the
    same random stream as a seeded CPU run of the reference;
  * `swav_args['batch_latents']` (default 1 = the reference) trains on several latents per
    optimiser step with joint-batch Sinkhorn; with torch.distributed initialised the batch
    is sharded over ranks and only the Sinkhorn marginals and the gradients are all-reduced;
  * only `projn_nw == 'linear'` (every shipped config); `sampling_method` 'random' (shipped) and 'patch'.
All random draws are made on the CPU generators (torch / numpy), a few KB per step.
"""
import os
import time

import numpy as np
import torch
import torch.nn as nn
from torchvision import transforms

from .. import _lib as L
from ..stylegan2.model import Generator
from . import engine as E


class SwAVClustering(object):

    def __init__(self, model, model_config, perturb_args, swav_args, sinkhorn_args, logger=None, train=True,
                 out_dir=None, device='cuda', tb=None, layer_hf_dim=None):
        L.load()  # fail loudly without the CUDA library / a B200
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("ganecdotes_b200.SwAVClustering has no CPU path (device must be 'cuda')")
        if not isinstance(model, Generator):
            model = Generator.from_reference(model, self.device)
        self.model = model.to(self.device)
        self.model_config = model_config
        self.perturb_args = perturb_args
        self.swav_args = swav_args
        self.writer = tb
        self.nclasses = swav_args['nclasses']
        self.nprototypes = swav_args['nprototypes']
        self.niters = sinkhorn_args['niters']
        self.eps = sinkhorn_args['eps']
        self.sinkhorn_args = sinkhorn_args.copy()
        self.logger = logger
        self.train = train
        self.out_dir = out_dir
        if out_dir is not None:
            self.swav_dir = os.path.join(self.out_dir, 'swav')
            os.makedirs(self.swav_dir, exist_ok=True)
            self.prototype_file = os.path.join(self.out_dir, 'prototypes.pt')
            self.projection_file = os.path.join(self.out_dir, 'projection.pt')
        else:
            self.prototype_file = self.projection_file = None
        if not self.train and self.projection_file is not None:
            if os.path.exists(self.projection_file):
                # whole pickled nn.Modules, like the reference (:84-86); torch >= 2.6 needs weights_only=False
                self.projection = torch.load(self.projection_file, weights_only=False).to(self.device)
                self.prototype = torch.load(self.prototype_file, weights_only=False).to(self.device)
            elif self.logger is not None:
                self.logger.info("Prototype File not found - pretraining ...")
        self.softmax_loss = nn.Softmax()
        with torch.no_grad():
            self.mean_latent = self._mean_latent(self.model_config.num_latents_for_mean)
            self.truncation = self.model_config.truncation
        self.fixed_transforms = transforms.Compose([
            transforms.RandomRotation(10),
            transforms.RandomHorizontalFlip(p=0.5),
        ])
        self.layer_hf_dim = layer_hf_dim
        self._head = None
        self._sk_ws = None
        # reproduce the reference's random stream (burn the draws of the work that is skipped)
        self.match_reference_rng = True
        self.passes_fwd = int(swav_args.get('passes_fwd', 3))
        self.passes_bwd = int(swav_args.get('passes_bwd', 1))
        self.proto_f16 = bool(swav_args.get('proto_f16', False))

    # ------------------------------------------------------------------ helpers
    def _mean_latent(self, n):
        z = torch.randn(n, self.model.style_dim)            # CPU generator, ref model.py:554-560
        return self.model.style(z.to(self.device)).mean(0, keepdim=True)

    def _burn_noise_draws(self, batch=1):
        """The reference's discarded forwards run with randomize_noise=True and draw one
        normal_() tensor per layer (ref model.py:380); keep the CPU stream aligned."""
        if not self.match_reference_rng:
            return
        for n in range(self.model.num_layers):
            res = 2 ** ((n + 5) // 2)
            torch.empty(batch, 1, res, res).normal_()

    def _dist_group(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return E.DistGroup(dist.group.WORLD, dist.get_rank(), dist.get_world_size())
        return None

    def _rows_per_patch(self):
        """rows one latent contributes per patch: patch_size pixels ('random') or a patch_size x patch_size
        crop ('patch', ref :150-158); None = every pixel."""
        ps = self.swav_args['patch_size']
        if ps is None or ps == self.model.size:
            return None
        return ps * ps if self.swav_args['sampling_method'] == 'patch' else ps

    def _step_config(self):
        return E.StepConfig(hlen=self.swav_args['hlen'], patch_size=self._rows_per_patch(),
                            num_patches=self.swav_args['num_patches'], niters=self.niters, eps=self.eps,
                            temperature=self.swav_args['temperature'], truncation=self.truncation,
                            perturb_std=list(self.perturb_args['perturb_std']),
                            source_pdf=self.sinkhorn_args.get('source_pdf', 'uniform'),
                            hf_interp=self._hf_interp())

    def _hf_interp(self):
        mode = self.swav_args.get('hf_interp', 'nearest')
        if mode not in ('nearest', 'bilinear'):
            raise NotImplementedError("hf_interp: 'nearest' (every shipped config) or 'bilinear'")
        if mode == 'bilinear' and self.sinkhorn_args.get('source_pdf', 'uniform') == 'image':
            raise NotImplementedError("hf_interp='bilinear' with source_pdf='image' (the norm image is gathered "
                                      "from nearest-upsampled rows)")
        return mode

    def _draw_view(self, b, layer_no):
        n_layers = self.perturb_args['n_layers']
        layers, pz = [], []
        for _ in range(b):
            self._burn_noise_draws(1)
            l = layer_no
            if l is None:
                l = np.random.choice(list(range(n_layers)))
            layers.append(int(l))
            pz.append(torch.cat([torch.randn(self.perturb_args['n_samples'], self.model.style_dim)
                                 for _ in range(2 * n_layers)], 0))
        return layers, torch.stack(pz)

    def draw_step(self, b_global):
        """All random draws of one optimiser step for `b_global` latents, in the
        reference's order."""
        d = self.model_config.latent_dim
        h = w = self.model.size
        z, vs, vt = [], [], []
        for _ in range(b_global):
            z.append(torch.randn(1, d))
            vs.append(self._draw_view(1, self.perturb_args['layer_no']))
            vt.append(self._draw_view(1, self.perturb_args['layer_no']))
        rot = []
        for _ in range(b_global):
            one = []
            for _v in range(2):
                ang = transforms.RandomRotation.get_params([-10.0, 10.0])
                flip = bool(torch.rand(1) < 0.5)
                one.append((ang, flip))
            rot.append(one)
        perms = []
        full = self.swav_args['patch_size'] is None or self.swav_args['patch_size'] == h
        by_patch = self.swav_args['sampling_method'] == 'patch'
        for _p in range(self.swav_args['num_patches']):
            if full:
                perms.append([torch.arange(h * w) for _ in range(b_global)])
            elif by_patch:      # ref :383-385: one offset per patch, used on both axes
                perms.append([E.patch_pick_rows(h, w, int(np.random.choice(h - self.swav_args['patch_size'])),
                                                self.swav_args['patch_size']) for _ in range(b_global)])
            else:
                perms.append([torch.randperm(h * w) for _ in range(b_global)])

        def mk(views, vi):
            return E.ViewDraws(layer_no=[v[0][0] for v in views], pert_z=torch.cat([v[1] for v in views], 0),
                               angle=[r[vi][0] for r in rot], flip=[r[vi][1] for r in rot])
        return E.StepDraws(z=torch.cat(z, 0), view_s=mk(vs, 0), view_t=mk(vt, 1), perms=perms)

    def lr_schedule(self, num_epochs, num_samples):
        """ref :303-317 (use_scheduler): linear warm-up from start_warmup to base_lr over warmup_epochs, then the
        reference's cosine to final_lr (its period is num_epochs - warmup_epochs *iterations*, as written there)."""
        import math
        a = self.swav_args
        warm = np.linspace(a['start_warmup'], a['base_lr'], num_samples * a['warmup_epochs'])
        iters = np.arange(num_samples * (a['num_epochs'] - a['warmup_epochs']))
        cos = np.array([a['final_lr'] + 0.5 * (a['base_lr'] - a['final_lr'])
                        * (1 + math.cos(math.pi * t / (a['num_epochs'] - a['warmup_epochs']))) for t in iters])
        return np.concatenate((warm, cos))

    @staticmethod
    def broadcast_draws(draws: E.StepDraws, group):
        """rank 0's draws of the global batch replace every other rank's (an object broadcast: for launchers that
        cannot re-seed; `pretrain` synchronises the generators' seed once instead)"""
        import torch.distributed as dist
        box = [draws if group.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group.pg)
        return box[0]

    @staticmethod
    def shard_draws(draws: E.StepDraws, rank, world):
        b = draws.z.shape[0]
        assert b % world == 0, "batch_latents must be divisible by the world size"
        per = b // world
        sl = slice(rank * per, (rank + 1) * per)

        def sv(v):
            return E.ViewDraws(v.layer_no[sl], v.pert_z[sl], v.angle[sl], v.flip[sl])
        return E.StepDraws(draws.z[sl], sv(draws.view_s), sv(draws.view_t), [p[sl] for p in draws.perms])

    def _features_nhwc(self, features):
        return [f.permute(0, 2, 3, 1).contiguous().float() for f in features]

    # ------------------------------------------------------------------ reference API
    def create_pixel_feature_vectors(self, features, pred=False):
        """ref :108-130 - materialises [B, hlen, H, W] (channels_last memory).  The training
        and prediction paths never call this: they gather straight into the GEMM operand."""
        feats = self._features_nhwc(features)
        h = max(f.shape[1] for f in feats)
        w = max(f.shape[2] for f in feats)
        if self.swav_args['hf_interp'] != 'nearest':
            raise NotImplementedError("create_pixel_feature_vectors materialises nearest-upsampled vectors only; "
                                      "hf_interp='bilinear' is applied inside the training / prediction paths")
        b = feats[0].shape[0]
        hlen = min(self.swav_args['hlen'], sum(f.shape[3] for f in feats))
        _, _, a = L.gather_rows(feats, h, w, hlen, None, None, b * h * w, want_lo=False, want_f32=True)
        return a.view(b, h, w, hlen).permute(0, 3, 1, 2)

    def _head_for_inference(self):
        return self.projection[0].weight.data

    def _proj_slope(self):
        """LeakyReLU slope of a '1-layer' projection network (ref :250-256), None for 'linear'; read from the module so
        that a `projection.pt` written by either implementation decides."""
        mods = list(self.projection)
        if len(mods) == 1:
            return None
        if len(mods) == 2 and isinstance(mods[1], nn.LeakyReLU):
            return float(mods[1].negative_slope)
        raise NotImplementedError("projection network: 'linear' or '1-layer' ('2-layer' puts BatchNorm1d over the "
                                  "sampled pixels of a patch between two Linear layers; no shipped config uses it)")

    def get_swav_codes_from_hidden_features(self, hfeat, new_shape=None, picks=None, train=False):
        """ref :133-182.  hfeat [1, D, H, W]; picks: a permutation ('random') or the crop offset ('patch')."""
        if self.swav_args['sampling_method'] not in ('random', 'patch'):
            raise NotImplementedError("sampling_method: 'random' or 'patch'")
        b, d, h, w = hfeat.shape
        x = hfeat.permute(0, 2, 3, 1).contiguous().float()
        row_src = row_img = None
        n = b * h * w
        if picks is not None:
            if self.swav_args['sampling_method'] == 'patch':
                picks = E.patch_pick_rows(h, w, int(picks), self.swav_args['patch_size'])
                n = picks.numel()
            else:
                n = self.swav_args['patch_size']
            row_src = picks[:n].to(torch.int32).to(self.device)
            row_img = torch.zeros(n, dtype=torch.int32, device=self.device)
        w_proj = self.projection[0].weight.data
        a_hi, a_lo, _ = L.gather_rows([x], h, w, d, row_img, row_src, n)
        wp_hi, wp_lo = L.split_planes(w_proj)
        z = L.gemm(a_hi, a_lo, wp_hi, wp_lo, n, w_proj.shape[0], d, 3)
        if self._proj_slope() is not None:
            z = E.proj_activation(z, self._proj_slope())
        if train:
            zn_hi, zn_lo, _ = L.l2norm_split(z)
            wk_hi, wk_lo = L.split_planes(self.prototype.weight.data)
            scores = L.gemm(zn_hi, zn_lo, wk_hi, wk_lo, n, self.prototype.weight.shape[0], z.shape[1], 3,
                            bias=self.prototype.bias.data)
        else:
            scores = z.t()
        if new_shape is not None:
            scores = scores.reshape(new_shape)
        return scores

    def preprocess(self, input_latent):
        if self.train:
            self.pretrain(input_latent)
        else:
            if self.projection_file is not None and os.path.exists(self.projection_file) and not self.train:
                pass
            else:
                self.pretrain(input_latent)

    def pretrain(self, input_latent, num_test_samples=5):
        """ref :205-505."""
        num_epochs = self.swav_args['num_epochs']
        num_samples = self.swav_args['num_samples']
        if self.swav_args['projn_nw'] not in ('linear', '1-layer'):
            raise NotImplementedError("projn_nw: 'linear' (every shipped config) or '1-layer'; '2-layer' puts "
                                      "BatchNorm1d over the sampled pixels of a patch between two Linear layers")
        if self.swav_args.get('add_local_loss', False):
            raise NotImplementedError("add_local_loss is broken in the reference (SURVEY §8 quirk 9) and off "
                                      "in every shipped config")
        if self.swav_args['sampling_method'] not in ('random', 'patch'):
            raise NotImplementedError("sampling_method: 'random' (every shipped config) or 'patch'")
        if int(self.perturb_args.get('n_samples', 1)) != 1:
            raise NotImplementedError("perturb_args['n_samples'] != 1: the reference would synthesise n_samples "
                                      "images per view (every shipped config uses 1)")
        ta = dict(self.swav_args['train_args'])
        unknown = set(ta) - {'lr', 'momentum', 'weight_decay'}
        if unknown or not ta.get('momentum', 0.0) >= 0:
            raise NotImplementedError(f"train_args {sorted(unknown)}: the fused LARC+SGD kernel implements lr, momentum "
                                      "and weight_decay (every shipped config: lr + momentum)")
        # test latents + their (unused) images: draws only (ref :222-238)
        for _ in range(num_test_samples):
            torch.randn(1, self.model_config.latent_dim)
        for _ in range(num_test_samples):
            self._burn_noise_draws(1)
        # same construction order / default init as the reference (CPU RNG), then to device
        layers = [nn.Linear(self.swav_args['hlen'], self.nclasses, bias=False)]
        if self.swav_args['projn_nw'] == '1-layer':               # ref :250-256
            layers.append(nn.LeakyReLU(inplace=True))
        self.projection = nn.Sequential(*layers).to(self.device)
        self.prototype = nn.Linear(self.nclasses, self.nprototypes).to(self.device)
        for p in list(self.projection.parameters()) + list(self.prototype.parameters()):
            p.requires_grad_(False)
        if self.logger is not None:
            self.logger.info("Projection Network:")
            self.logger.info(self.projection.__str__())
            self.logger.info("Prototype Matrix:")
            self.logger.info(self.prototype.__str__())
        group = self._dist_group()
        world = group.world if group is not None else 1
        if group is not None:
            # replicas must start from identical weights / mean latent whatever each rank's CPU RNG state is
            import torch.distributed as dist
            for t in (self.projection[0].weight.data, self.prototype.weight.data, self.prototype.bias.data,
                      self.mean_latent):
                dist.broadcast(t, src=0, group=group.pg)
            # every rank draws the whole global batch from its own CPU generators and keeps its shard: the generators
            # are re-seeded from one number drawn on rank 0, so the streams agree whatever the launcher seeded
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64).to(self.device)
            dist.broadcast(seed, src=0, group=group.pg)
            torch.manual_seed(int(seed.item()))
            np.random.seed(int(seed.item()) % (2 ** 32))
        self._head = E.SwavHead(self.projection[0].weight.data, self.prototype.weight.data,
                                self.prototype.bias.data, ta['lr'], ta.get('momentum', 0.0),
                                self.swav_args['trust_coeff'], self.passes_fwd, self.passes_bwd, self.proto_f16,
                                weight_decay=ta.get('weight_decay', 0.0), proj_slope=self._proj_slope())
        self._sk_ws = L.SinkhornWorkspace(self.nprototypes, self.device)
        lr_schedule = self.lr_schedule(num_epochs, num_samples) if self.swav_args.get('use_scheduler', False) else None
        b_global = int(self.swav_args.get('batch_latents', 1))
        cfg = self._step_config()
        t0 = time.time()
        loss = None
        # The host draws of a step do not depend on device results: the inputs of step i+1 are drawn and
        # uploaded on a side stream while the compute stream runs step i.
        side = torch.cuda.Stream(device=self.device)

        def stage_next():
            draws = self.draw_step(b_global)
            if group is not None:
                draws = self.shard_draws(draws, group.rank, world)
            return E.prepare_step_inputs(self.model, draws, cfg, self.device, stream=side)

        total = num_epochs * num_samples
        nxt = stage_next() if total > 0 else None
        done = 0
        for e in range(num_epochs):
            for i in range(num_samples):
                inp, done = nxt, done + 1
                if lr_schedule is not None:                      # ref :323-326
                    self._head.lr = float(lr_schedule[e * num_samples + i])
                loss = E.swav_train_step_device(self.model, self._head, self.mean_latent, inp, cfg, group,
                                                self._sk_ws)
                nxt = stage_next() if done < total else None
                if self.writer is not None:
                    self.writer.add_scalar('swav/loss', loss, e)
            if self.logger is not None and e % self.swav_args['epoch_print_freq'] == 0:
                self.logger.info(f" E:{e}\t|\tLoss: {float(loss):.03f} \t|\tT: {time.time() - t0:.03f}")
        if self.logger is not None:
            self.logger.info("Finished pretraining - Saving projection file")
        if self.prototype_file is not None and (group is None or group.rank == 0):
            torch.save(self.prototype, self.prototype_file)
            torch.save(self.projection, self.projection_file)

    def sinkhorn_knopp(self, scores, img):
        """ref :509-544.  scores [N,K] -> Q [N,K] (materialised for API parity)."""
        s = scores.detach().float().contiguous()
        n, k = s.shape
        r = c = None
        if self.sinkhorn_args['source_pdf'] == 'image':
            histb = torch.histc(img, n) + 1e-9
            histb[0] = histb[1]
            c = (histb / histb.sum()).float().contiguous()
            histk = torch.histc(img, k) + 1e-9
            histk[0] = histk[1]
            r = (histk / histk.sum()).float().contiguous()
        ws = L.SinkhornWorkspace(k, s.device)
        la = E.sinkhorn_log_a(s, self.niters, self.eps, ws, n, None, r, c)
        return L.sinkhorn_q(s, 1.0 / self.eps, la)

    def calculate_swapped_prediction_loss(self, softmax_ns, softmax_nt, code_ns, code_nt):
        """ref :547-570, for already materialised codes.  (The training loop uses the fused
        loss+gradient kernel instead; this method exists for API parity.)"""
        lst = torch.mean(torch.sum(code_ns * torch.log_softmax(softmax_nt, dim=1), dim=1))
        lts = torch.mean(torch.sum(code_nt * torch.log_softmax(softmax_ns, dim=1), dim=1))
        return -0.5 * (lst + lts)

    def create_hidden_features_from_perturbed_vectors(self, layer_no=None, input_latent=None, input_is_latent=True):
        """ref :574-656.  Returns (hfeat [1,hlen,H,W], perturbed_img, layer_no)."""
        with torch.no_grad():
            if input_latent is None:
                input_latent = torch.randn(1, self.model_config.latent_dim).to(self.device)
                input_is_latent = False
            w = input_latent.to(self.device).float()
            if not input_is_latent:
                w = self.model.style(w)
            layers, pz = self._draw_view(1, layer_no)
            view = E.ViewDraws(layer_no=layers, pert_z=pz, angle=[0.0], flip=[False])
            wplus = E.view_wplus(self.model, w, self.mean_latent, self.truncation, view,
                                 self.perturb_args['perturb_std'])
            img, feats = self.model.synthesize(wplus, None, need_image=True)
            hfeat = self.create_pixel_feature_vectors([f.permute(0, 3, 1, 2) for f in feats])
        return hfeat, img, layers[0]

    def predict_swav_codes(self, input_latent, input_is_latent=True):
        """ref :659-693 (the reference ignores `input_is_latent`: always a W latent)."""
        return E.predict_codes(self.model, self.projection[0].weight.data, input_latent, self.mean_latent,
                               self.model_config.truncation, self.swav_args['hlen'], self.passes_fwd,
                               hf_interp=self._hf_interp(), proj_slope=self._proj_slope())
