"""Drop-in for the reference's k-means preprocessor `baseline/hfc_kmeans/segmentor.py::HFCPreprocessor`
(:11-226; the plugin socket `baseline/hfc_kmeans/base.py` re-exports it as `preprocessor`): fits one flat k-means
per StyleGAN block on the hidden features of latent-perturbed samples of the one-shot image and turns the hidden
features of a latent into per-pixel one-hot cluster maps.

Runs on the sm_100a path: the synthesis network, the `lib/oneshot` perturbation functions
(`ganecdotes_b200.oneshot`), the Lloyd fit and the nearest-centre assignment (`FlatKMeansAssign`: `gx_kmeans_assign`,
tensor-core scores for large inputs, `gx_onehot_nearest`).  `hfc_algo='hfc_kmeans'` with `hier_encode=False` - the
shipped `hfc_kmeans_config.py`; the hierarchical variants raise.  Artefacts: `kmeans_centers.pt` (a list of [K_n, C_n]
centre tensors) in `out_dir`; scikit-learn pickles written by the reference (`clusterer_layer_n.sav`) are read when
present and scikit-learn is importable.
"""
import os
import pickle

import torch

from .. import oneshot
from ..stylegan2.model import Generator
from .hfc_kmeans_clustering import FlatKMeansAssign


class HFCPreprocessor(object):

    def __init__(self, model, model_config, perturb_args, hfc_args, hfc_algo='hfc_kmeans', hier_encode=True,
                 hle_samples=500, train=True, out_dir=None, logger=None):
        if hfc_algo != 'hfc_kmeans':
            raise NotImplementedError("hfc_algo: 'hfc_kmeans' (the shipped config); 'hfc_kmeans_hier' is not built")
        if hier_encode:
            raise NotImplementedError("hier_encode=True (Bayesian belief encoding) is not built; the shipped "
                                      "hfc_kmeans_config.py sets hier_encode=False")
        if not isinstance(model, Generator):
            model = Generator.from_reference(model, 'cuda')
        self.model = model
        self.perturb_config = perturb_args
        self.hfc_args = hfc_args
        self.hier_encode = hier_encode
        self.hfc_algo = hfc_algo
        self.out_dir = out_dir
        self.train = train
        self.logger = logger
        self.model_config = model_config
        self.hle_samples = hle_samples
        base = hfc_args['base_args']
        self.n_layers = base['n_layers']
        self.clusters_per_layer = list(base['clusters_per_layer'])
        self.out_size = base['out_size']
        self.kmeans_args = dict(hfc_args.get('kmeans_args', {}))
        self.hfc_model = None
        self.centers_file = os.path.join(out_dir, 'kmeans_centers.pt') if out_dir is not None else None
        if base.get('presaved', False) or not train:
            self._load()

    # ------------------------------------------------------------------ artefacts
    def _load(self):
        dev = next(self.model.parameters()).device
        if self.centers_file is not None and os.path.exists(self.centers_file):
            self.hfc_model = FlatKMeansAssign(torch.load(self.centers_file), self.out_size, dev)
            return
        savs = [os.path.join(self.out_dir or '.', f"clusterer_layer_{n}.sav") for n in range(self.n_layers)]
        if all(os.path.exists(p) for p in savs):         # the reference's pickled sklearn.cluster.KMeans models
            self.hfc_model = FlatKMeansAssign.from_sklearn([pickle.load(open(p, 'rb')) for p in savs], self.out_size, dev)
            return
        if not self.train:
            raise FileNotFoundError('Models not found - use train_hfc_model() to create the model first!')

    def _w_plus(self, input_latent, mean_latent, truncation):
        """`self.model([latent], return_latents=True, truncation..., input_is_latent=True)[1]` (ref :92-97)"""
        dev = next(self.model.parameters()).device
        lat = input_latent.to(dev).float()
        if lat.dim() == 2 and lat.shape[0] > 1:
            lat = lat.unsqueeze(0)
        _, w_latents = self.model([lat], return_latents=True, truncation_latent=mean_latent, truncation=truncation,
                                  input_is_latent=True)
        return w_latents.detach().clone()

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def train_hfc_model(self, input_latent, return_aug=False):
        """ref :68-165: for every block k the two W+ rows of the block are perturbed for `n_samples` samples, the
        block's hidden features ([n_samples, C_k, h_k, w_k], the two maps of the resolution concatenated) are
        clustered with K_k centres."""
        pc = self.perturb_config
        mean_latent = self.model.mean_latent(self.model_config.num_latents_for_mean)
        truncation = pc['truncation']
        w = self._w_plus(input_latent, mean_latent, truncation)
        hidden, new_latent_list = [], []
        for k in range(pc['n_layers']):
            stds = [0] * (2 * pc['n_layers'])
            stds[2 * k] = stds[2 * k + 1] = pc['perturb_std'][k]
            pl = oneshot.create_perturbed_vectors_from_latents(w, self.model, n_samples=pc['n_samples'],
                                                               n_layers=pc['n_layers'], perturb_std=stds)
            new = w.repeat(pc['n_samples'], 1, 1)
            new[:, 2 * k, :], new[:, 2 * k + 1, :] = pl[2 * k], pl[2 * k + 1]
            new_latent_list += [pl[2 * k], pl[2 * k + 1]]
            _, hfeat = oneshot.create_images_and_features_from_perturbed_latents(
                new, self.model, {'truncation': truncation, 'mean_latent': mean_latent}, layer_no=k,
                return_feat=True, return_image=True, skip_const=True)
            hidden.append(hfeat)
            if self.logger is not None:
                self.logger.info(f"Generated features for Layer: {k}")
        self.hfc_model = FlatKMeansAssign.fit_grouped(hidden, self.clusters_per_layer, self.out_size,
                                                      seed=int(self.kmeans_args.get('random_state') or 0),
                                                      max_iter=int(self.kmeans_args.get('max_iter', 300)),
                                                      tol=float(self.kmeans_args.get('tol', 1e-4)))
        if self.centers_file is not None:
            torch.save([c.cpu() for c in self.hfc_model.centers], self.centers_file)
        if self.logger is not None:
            for n in range(self.n_layers):
                self.logger.info(f"Fitted model for Layer {n}")
        if return_aug:
            return hidden, new_latent_list

    @torch.no_grad()
    def predict_hfc_vectors(self, input_latent):
        """ref :168-226 with hier_encode=False: per-pixel cluster maps in {-1, +1} [B, sum K, out, out] and the list of
        per-layer label maps [B, 1, h_n, w_n]."""
        if self.hfc_model is None:
            self._load()
        if self.hfc_model is None:
            raise FileNotFoundError('Models not found - use train_hfc_model() to create the model first!')
        mean_latent = self.model.mean_latent(self.model_config.num_latents_for_mean)
        truncation = self.perturb_config['truncation']
        w = self._w_plus(input_latent, mean_latent, truncation)
        _, hfeat = oneshot.create_images_and_features_from_perturbed_latents(
            w, self.model, {'truncation': 0.7, 'mean_latent': mean_latent}, return_feat=True, return_image=True,
            skip_const=True)                                       # the literal 0.7 is the reference's (:198)
        return self.hfc_model.predict_grouped(hfeat[:self.perturb_config['n_layers']], pm1=True)
