"""Drop-in for the reference's SimCLR baseline `baseline/hfc_with_simclr/simclr_clustering.py::SimCLRClustering`
(same constructor, `pretrain`, `predict_simclr_codes`, `projection.pt` artefact) on the sm_100a kernels.

Head: Linear(hlen -> C, no bias) -> BatchNorm1d -> LeakyReLU(0.01) -> Linear(C -> C, no bias) on channel-normalised
per-pixel vectors (ref :150-160, :192, :202).  Training batch: `batch_size` random pixels of each of the two views,
interleaved s_0, t_0, s_1, ... (ref :215-226), contrastive loss exactly as the reference's two O(n^2) Python loops
evaluate it (both quirks kept, see `gx_simclr_loss`), LARC + SGD(momentum).

What runs where: the synthesis network and the gather of the sampled rows are the SwAV path's kernels; the two
Linear layers and their three gradient GEMMs are `gx_gemm` (3-pass split-bf16); BatchNorm / LeakyReLU forward and
backward, the loss + its gradient and the optimiser are the kernels of csrc/gx_simclr.cu and gx_head.cu.  F.normalize
is folded in: W1 (f / |f|) = (W1 f) / |f|, so the first GEMM runs on the un-normalised rows (for prediction: on
every level at its native resolution, `engine.project_all_pixels`) and 1 / |f| is a row scale inside the BatchNorm
kernels.  The random draws are made on the CPU generators in the reference's order (per view: noise buffers of the
discarded forward, layer choice, 2 * n_layers perturbation draws, rotation angle, flip; then one randperm).
"""
import os
import time
from dataclasses import dataclass
from typing import List

import numpy as np
import torch
import torch.nn as nn
from torchvision import transforms

from .. import _lib as L
from ..hfc_with_swav import engine as E
from ..stylegan2.model import Generator

BN_EPS = 1e-5
LRELU_SLOPE = 0.01          # nn.LeakyReLU() default, ref :155


@dataclass
class SimCLRDraws:
    """random draws of one iteration (ref :175-214)"""
    z: torch.Tensor                     # [1, D]
    layer_no: List[int]                 # [s, t]
    pert_z: torch.Tensor                # [2, 2*n_layers, D]
    angle: List[float]
    flip: List[bool]
    perm: torch.Tensor                  # randperm(H*W), shared by both views


class SimCLRHead:
    """parameters of the projection network (in place on the nn.Modules the caller saves), momentum buffers"""

    def __init__(self, projection: nn.Sequential, lr, momentum, trust, weight_decay=0.0):
        self.lin1, self.bn, self.lin2 = projection[0], projection[1], projection[3]
        self.params = [self.lin1.weight.data, self.bn.weight.data, self.bn.bias.data, self.lin2.weight.data]
        self.bufs = [torch.zeros_like(p) for p in self.params]
        self.lr, self.momentum, self.trust, self.weight_decay = lr, momentum, trust, float(weight_decay)
        self.norms = L.larc_scratch(self.params[0].device)
        self.steps = 0

    def optimizer_step(self, grads):
        first = 1 if self.steps == 0 else 0
        for p, g, m in zip(self.params, grads, self.bufs):
            L.larc_sgd_(p, g.contiguous(), m, self.lr, self.momentum, self.trust, self.weight_decay, 1e-8, first, self.norms)
        self.steps += 1


@torch.no_grad()
def simclr_train_step(gen, head: SimCLRHead, mean_latent, draws: SimCLRDraws, hlen, batch_size, temperature,
                      truncation, perturb_std):
    """One iteration of SimCLRClustering.pretrain (ref :175-273).  Returns (loss tensor [1], grads list)."""
    dev = head.params[0].device
    h = w = gen.size
    c = head.lin1.weight.shape[0]
    n2 = 2 * batch_size
    wlat = gen.style(draws.z.to(dev).float())
    view = E.ViewDraws(layer_no=list(draws.layer_no) * 1, pert_z=draws.pert_z, angle=draws.angle, flip=draws.flip)
    # both views as one batch of 2 images (same latent, two perturbed W+)
    wplus = E.view_wplus(gen, wlat.repeat(2, 1), mean_latent, truncation, view, perturb_std)
    _, feats = gen.synthesize(wplus, None, need_image=False)
    # rows interleaved s_0, t_0, s_1, t_1, ...: image index alternates, both views share the sampled pixels
    picks = draws.perm[:batch_size]
    src = torch.stack([E.rotate_flip_index_map(h, w, draws.angle[v], draws.flip[v])[picks] for v in range(2)], 1)
    row_src = src.reshape(-1).to(torch.int32).to(dev)
    row_img = torch.tensor([0, 1] * batch_size, dtype=torch.int32, device=dev)
    a_hi, a_lo, _, nrm = L.gather_rows(feats, h, w, hlen, row_img, row_src, n2, want_lo=True, want_norm=True)
    rscale = L.recip_clamp(nrm, 1e-12)                                   # F.normalize(dim=1), ref :192, :202
    w1, gamma, beta, w2 = head.params
    w1_hi, w1_lo = L.split_planes(w1)
    hraw = L.gemm(a_hi, a_lo, w1_hi, w1_lo, n2, c, hlen, 3, tag="gemm_simclr")
    mean, invstd = L.bn_stats(hraw, rscale, BN_EPS, head.bn.running_mean, head.bn.running_var, head.bn.momentum or 0.1)
    a1, a1_hi, a1_lo = L.bn_act_apply(hraw, rscale, mean, invstd, gamma, beta, LRELU_SLOPE)
    w2_hi, w2_lo = L.split_planes(w2)
    z = L.gemm(a1_hi, a1_lo, w2_hi, w2_lo, n2, w2.shape[0], c, 3, tag="gemm_simclr")
    loss, dz = L.simclr_loss(z, temperature)
    # backward: dW2 = dz^T a1, da1 = dz W2, through lrelu + BN, dW1 = dhraw^T a
    dz_hi, dz_lo = L.split_planes(dz)
    g_w2 = torch.zeros_like(w2)
    L.gemm(dz_hi, dz_lo, a1_hi, a1_lo, w2.shape[0], c, n2, 3, out=g_w2, a_mn=True, b_mn=True, accumulate=True,
           tag="gemm_simclr")
    w2t_hi, w2t_lo = L.split_planes(w2, transpose=True)
    da1 = L.gemm(dz_hi, dz_lo, w2t_hi, w2t_lo, n2, c, w2.shape[0], 3, tag="gemm_simclr")
    dhs, g_gamma, g_beta = L.bn_act_bwd(da1, hraw, rscale, mean, invstd, gamma, beta, LRELU_SLOPE)
    dh_hi, dh_lo = L.split_planes(dhs)
    g_w1 = torch.zeros_like(w1)
    L.gemm(dh_hi, dh_lo, a_hi, a_lo, c, hlen, n2, 3, out=g_w1, a_mn=True, b_mn=True, accumulate=True, tag="gemm_simclr")
    grads = [g_w1, g_gamma, g_beta, g_w2]
    head.bn.num_batches_tracked += 1
    head.optimizer_step(grads)
    return loss, grads


class SimCLRClustering(object):

    def __init__(self, model, model_config, perturb_args, simclr_args, logger=None, train=True, out_dir=None,
                 device='cuda', tb=None, layer_hf_dim=None):
        L.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("ganecdotes_b200.SimCLRClustering has no CPU path (device must be 'cuda')")
        if not isinstance(model, Generator):
            model = Generator.from_reference(model, self.device)
        self.model = model.to(self.device)
        self.model_config = model_config
        self.perturb_args = perturb_args
        self.simclr_args = simclr_args
        self.writer = tb
        self.nclasses = simclr_args['nclasses']
        self.logger = logger
        self.train = train
        self.out_dir = out_dir
        self.projection_file = None
        if out_dir is not None:
            self.swav_dir = os.path.join(self.out_dir, 'simclr')
            os.makedirs(self.swav_dir, exist_ok=True)
            self.projection_file = os.path.join(self.out_dir, 'projection.pt')
            if not self.train:
                if os.path.exists(self.projection_file):
                    self.projection = torch.load(self.projection_file, weights_only=False).to(self.device)
                elif self.logger is not None:
                    self.logger.info("Projection File not found - pretraining ...")
        with torch.no_grad():
            z = torch.randn(self.model_config.num_latents_for_mean, self.model.style_dim)       # CPU generator
            self.mean_latent = self.model.style(z.to(self.device)).mean(0, keepdim=True)
            self.truncation = self.model_config.truncation
        self.fixed_transforms = transforms.Compose([transforms.RandomRotation(10),
                                                    transforms.RandomHorizontalFlip(p=0.5)])
        self.layer_hf_dim = layer_hf_dim
        self.match_reference_rng = True
        self._head = None

    # ------------------------------------------------------------------ helpers
    def _burn_noise_draws(self):
        """the reference's discarded forward draws one normal_() tensor per layer (ref model.py:380)"""
        if not self.match_reference_rng:
            return
        for n in range(self.model.num_layers):
            res = 2 ** ((n + 5) // 2)
            torch.empty(1, 1, res, res).normal_()

    def draw_step(self) -> SimCLRDraws:
        d = self.model_config.latent_dim
        n_layers = self.perturb_args['n_layers']
        z = torch.randn(1, d)
        layers, pz, ang, flip = [], [], [], []
        for _v in range(2):                                   # ref :187-203: each view draws its own transform
            self._burn_noise_draws()
            l = self.perturb_args['layer_no']
            if l is None:
                l = np.random.choice(list(range(n_layers)))
            layers.append(int(l))
            pz.append(torch.cat([torch.randn(self.perturb_args['n_samples'], d) for _ in range(2 * n_layers)], 0))
            ang.append(transforms.RandomRotation.get_params([-10.0, 10.0]))
            flip.append(bool(torch.rand(1) < 0.5))
        h = self.model.size
        return SimCLRDraws(z=z, layer_no=layers, pert_z=torch.stack(pz), angle=ang, flip=flip,
                           perm=torch.randperm(h * h))

    def preprocess(self, input_latent):
        if self.train or self.projection_file is None or not os.path.exists(self.projection_file):
            self.pretrain(input_latent)

    # ------------------------------------------------------------------ reference API
    def create_pixel_feature_vectors(self, features, pred=False):
        """ref :90-112 - materialises [B, hlen, H, W]; the training / prediction paths never call it"""
        feats = [f.permute(0, 2, 3, 1).contiguous().float() for f in features]
        h = max(f.shape[1] for f in feats)
        w = max(f.shape[2] for f in feats)
        if self.simclr_args.get('hf_interp', 'nearest') != 'nearest':
            raise NotImplementedError("hf_interp: 'nearest' (the shipped SimCLR config)")
        b = feats[0].shape[0]
        hlen = min(self.simclr_args['hlen'], sum(f.shape[3] for f in feats))
        _, _, a = L.gather_rows(feats, h, w, hlen, None, None, b * h * w, want_lo=False, want_f32=True)
        return a.view(b, h, w, hlen).permute(0, 3, 1, 2)

    def pretrain(self, input_latent, num_test_samples=2):
        """ref :131-281"""
        a = self.simclr_args
        if int(self.perturb_args.get('n_samples', 1)) != 1:
            raise NotImplementedError("perturb_args['n_samples'] != 1")
        ta = dict(a['train_args'])
        if set(ta) - {'lr', 'momentum', 'weight_decay'}:
            raise NotImplementedError("train_args: lr, momentum, weight_decay")
        self.projection = nn.Sequential(nn.Linear(a['hlen'], self.nclasses, bias=False), nn.BatchNorm1d(self.nclasses),
                                        nn.LeakyReLU(inplace=True),
                                        nn.Linear(self.nclasses, self.nclasses, bias=False)).to(self.device)
        for p in self.projection.parameters():
            p.requires_grad_(False)
        self._head = SimCLRHead(self.projection, ta['lr'], ta.get('momentum', 0.0), a['trust_coeff'],
                                ta.get('weight_decay', 0.0))
        t0 = time.time()
        loss = None
        for e in range(a['num_iters']):
            draws = self.draw_step()
            loss, _ = simclr_train_step(self.model, self._head, self.mean_latent, draws, a['hlen'], a['batch_size'],
                                        a['temperature'], self.truncation, list(self.perturb_args['perturb_std']))
            if self.writer is not None:
                self.writer.add_scalar('simclr/loss', float(loss), e)
            if self.logger is not None:
                self.logger.info(f" (Iter:{e}):\tLoss: {float(loss):.03f},\tTime: {time.time() - t0:.03f}\t")
        if self.projection_file is not None:
            torch.save(self.projection, self.projection_file)

    @torch.no_grad()
    def predict_simclr_codes(self, input_latent, images_per_chunk=4):
        """ref :362-401: codes [B, C, H, W] (channels_last memory) and the first arg-max label map.  BatchNorm uses
        batch statistics when the module is in training mode (what the reference does right after `pretrain`) and the
        running statistics in eval mode (a projection loaded with `train=False` then `.eval()`)."""
        proj = self.projection
        lin1, bn, lin2 = proj[0], proj[1], proj[3]
        dev = self.device
        w = input_latent.to(dev).float()
        mean = self.mean_latent.reshape(-1).float().contiguous()
        wt = L.truncate(w.contiguous(), mean, self.model_config.truncation) if self.model_config.truncation < 1 else w
        latent = wt.unsqueeze(1).expand(-1, self.model.n_latent, -1) if wt.dim() == 2 else wt
        _, feats = self.model.synthesize(latent, None, need_image=False)
        b, h = latent.shape[0], self.model.size
        hlen, c = self.simclr_args['hlen'], lin1.weight.shape[0]
        w1_hi, w1_lo = L.split_planes(lin1.weight.data.contiguous())
        w2_hi, w2_lo = L.split_planes(lin2.weight.data.contiguous())
        cout = lin2.weight.shape[0]
        codes = torch.empty((b * h * h, cout), dtype=torch.float32, device=dev)
        labels = torch.empty((b * h * h,), dtype=torch.int64, device=dev)
        if bn.training and b > 1:
            images_per_chunk = b            # batch statistics are over all rows of the call
        for i0 in range(0, b, images_per_chunk):
            i1 = min(b, i0 + images_per_chunk)
            sub = [f[i0:i1] for f in feats]
            n = (i1 - i0) * h * h
            hraw, _ = E.project_all_pixels(w1_hi, w1_lo, sub, i1 - i0, h, h, hlen, 3)
            _, _, _, nrm = L.gather_rows(sub, h, h, hlen, None, None, n, want_planes=False, want_norm=True)
            rscale = L.recip_clamp(nrm, 1e-12)
            if bn.training:
                bmean, invstd = L.bn_stats(hraw, rscale, bn.eps)
            else:
                bmean = bn.running_mean.float().contiguous()
                invstd = L.rsqrt_eps(bn.running_var.float().contiguous(), bn.eps)
            _, a_hi, a_lo = L.bn_act_apply(hraw, rscale, bmean, invstd, bn.weight.data, bn.bias.data, LRELU_SLOPE,
                                           want_f32=False)
            zc = codes[i0 * h * h: i1 * h * h]
            L.gemm(a_hi, a_lo, w2_hi, w2_lo, n, cout, c, 3, out=zc, tag="gemm_simclr", pair=True)
            L.argmax_rows(zc, out=labels[i0 * h * h: i1 * h * h])
        return codes.view(b, h, h, cout).permute(0, 3, 1, 2), labels.view(b, h, h)
