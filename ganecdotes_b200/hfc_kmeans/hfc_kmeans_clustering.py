"""Per-pixel k-means assignment on GAN features (BASELINE config 5) - the `predict` side of the
reference's `FlatKMeansHFC` (baseline/hfc_kmeans/hfc_kmeans_clustering.py:169-230).

For each of the `n_layers` resolutions, the two same-resolution feature maps f[2n+1], f[2n+2]
(ref lib/oneshot/image_augmentor.py:80-90 with skip_const=True) form the per-pixel vector; every
pixel is assigned to its nearest cluster centre (first index on ties, like sklearn's argmin);
labels become one-hot maps resized NEAREST to `out_size` and concatenated over layers.
`kmeans_fit` is the `fit` side (ref :146-166, sklearn.cluster.KMeans with its defaults: greedy k-means++
seeding, Lloyd iterations, tol relative to the mean feature variance) on the same assignment kernel;
scikit-learn's result depends on its version and RNG, so parity is statistical (inertia), SURVEY §8(f) rank 2.
Fitted scikit-learn centres can be passed as well (`from_sklearn`).
"""
from typing import List, Sequence

import torch

from .. import _lib as L


@torch.no_grad()
def kmeans_fit(x, k, max_iter=300, tol=1e-4, seed=0, x2=None):
    """Lloyd k-means on x [n,c] (+ x2 [n,c2], the second map of the resolution) fp32 CUDA tensors.
    Returns (centers [k,c(+c2)], labels int32 [n], inertia float, n_iter).

    Seeding: greedy k-means++ (2 + log k candidates per step, drawn with probability proportional to the
    squared distance to the nearest chosen centre - scikit-learn's `_kmeans_plusplus`).  Each iteration is one
    launch of the assignment kernel (labels + squared distances) and a deterministic segment mean
    (`gx_segment_sum_rows` over the label-sorted rows).  Converged when the squared centre shift is below
    tol * mean(var(x, axis=0)), like scikit-learn."""
    import math
    xs = [x.float().contiguous()] + ([x2.float().contiguous()] if x2 is not None else [])
    n = xs[0].shape[0]
    dev = xs[0].device
    gen = torch.Generator(device=dev).manual_seed(int(seed))

    def row(i):
        return torch.cat([t[i] for t in xs]).unsqueeze(0).contiguous()

    def dist_to(c):
        return L.kmeans_assign(xs[0], c, xs[1] if len(xs) > 1 else None, want_dist=True)

    trials = 2 + int(math.log(k))
    first = int(torch.randint(n, (1,), generator=gen, device=dev))
    centers = [row(first)]
    _, closest = dist_to(centers[0])
    for _ in range(1, k):
        cand = torch.multinomial(closest.clamp_min(0) + 1e-30, trials, replacement=True, generator=gen)
        best = None
        for j in cand.tolist():
            _, d = dist_to(row(j))
            d = torch.minimum(d, closest)
            pot = float(d.sum())
            if best is None or pot < best[0]:
                best = (pot, j, d)
        centers.append(row(best[1]))
        closest = best[2]
    c = torch.cat(centers).contiguous()
    var = torch.cat([t.var(dim=0, unbiased=False) for t in xs]).mean().item()
    labels = inertia = None
    it = 0
    for it in range(1, max_iter + 1):
        labels, d = L.kmeans_assign(xs[0], c, xs[1] if len(xs) > 1 else None, want_dist=True)
        inertia = float(d.sum())
        order = torch.argsort(labels.long(), stable=True).to(torch.int32)
        counts = torch.bincount(labels.long(), minlength=k)
        seg_off = torch.zeros(k + 1, dtype=torch.int32, device=dev)
        seg_off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        sums = torch.cat([L.segment_sum_rows(t, order, seg_off, k, want_planes=False, want_f32=True)[2] for t in xs], 1)
        new_c = sums / counts.clamp_min(1).unsqueeze(1).float()
        empty = counts == 0
        if empty.any():      # relocate empty clusters to the points farthest from their centre
            far = torch.topk(d, int(empty.sum())).indices
            new_c[empty] = torch.cat([row(int(i)) for i in far])
        shift = float(((new_c - c) ** 2).sum())
        c = new_c.contiguous()
        if shift <= tol * var:
            break
    labels, d = L.kmeans_assign(xs[0], c, xs[1] if len(xs) > 1 else None, want_dist=True)
    return c, labels, float(d.sum()), it


class FlatKMeansAssign(object):
    def __init__(self, centers: Sequence[torch.Tensor], out_size: int = 256, device="cuda"):
        """centers[n]: [K_n, C_n] float tensor or numpy array (e.g. `KMeans.cluster_centers_`)."""
        self.device = torch.device(device)
        self.centers = [torch.as_tensor(c, dtype=torch.float32).contiguous().to(self.device) for c in centers]
        self.clusters_per_layer = [c.shape[0] for c in self.centers]
        self.n_layers = len(self.centers)
        self.out_size = out_size

    @classmethod
    def from_sklearn(cls, clusterers, out_size=256, device="cuda"):
        return cls([c.cluster_centers_ for c in clusterers], out_size, device)

    @classmethod
    def fit(cls, features, clusters_per_layer, out_size=256, seed=0, **kmeans_args):
        """ref `_layerwise_fit` (:146-166) for every layer: features = the generator's feature list
        ([B,C,H,W] channels_last views); layer n clusters the per-pixel vectors of maps 2n+1, 2n+2."""
        centers = []
        for n, k in enumerate(clusters_per_layer):
            f1 = features[2 * n + 1].permute(0, 2, 3, 1).contiguous().float()
            f2 = features[2 * n + 2].permute(0, 2, 3, 1).contiguous().float()
            c, _, _, _ = kmeans_fit(f1.reshape(-1, f1.shape[3]), k, seed=seed + n, x2=f2.reshape(-1, f2.shape[3]),
                                    **kmeans_args)
            centers.append(c)
        return cls(centers, out_size, features[0].device)

    @classmethod
    def fit_grouped(cls, grouped, clusters_per_layer, out_size=256, seed=0, **kmeans_args):
        """`fit` on already grouped per-layer features: grouped[n] = [B, C_n, h, w] (the two maps of block n
        concatenated, what `create_images_and_features_from_perturbed_latents(..., skip_const=True)` returns)"""
        centers = []
        for n, k in enumerate(clusters_per_layer):
            f = grouped[n].permute(0, 2, 3, 1).contiguous().float()
            c, _, _, _ = kmeans_fit(f.reshape(-1, f.shape[3]), k, seed=seed + n, **kmeans_args)
            centers.append(c)
        return cls(centers, out_size, grouped[0].device)

    def predict_grouped(self, grouped, pm1=False):
        """ref BaseHFCModel.predict (:93-110) on grouped per-layer features [B, C_n, h, w]:
        (one-hot maps [B, sum K, out, out], [labels [B, 1, h_n, w_n]]); pm1: maps in {-1, +1} (the `* 2 - 1` of
        predict_hfc_vectors, ref segmentor.py:222-226, written by the same kernel)"""
        labs = []
        b = grouped[0].shape[0]
        maps = torch.empty((b, sum(self.clusters_per_layer), self.out_size, self.out_size), dtype=torch.float32,
                           device=self.device)
        off = 0
        for n in range(self.n_layers):
            f = grouped[n].permute(0, 2, 3, 1).contiguous().float()
            lab, _ = self._layerwise_predict([f], n, out=maps[:, off:off + self.clusters_per_layer[n]], pm1=pm1)
            off += self.clusters_per_layer[n]
            labs.append(lab)
        return maps, labs

    def _layerwise_predict(self, feats_nhwc: List[torch.Tensor], n: int, out=None, pm1=False):
        """ref :169-208.  feats_nhwc: the map(s) of layer n as fp32 NHWC tensors.
        Returns (labels int32 [b,1,h,w], label_maps float [b,K,out,out]); `out`: where to write the maps."""
        f1 = feats_nhwc[0]
        b, h, w, c1 = f1.shape
        x1 = f1.reshape(-1, c1)
        x2 = feats_nhwc[1].reshape(-1, feats_nhwc[1].shape[3]) if len(feats_nhwc) > 1 else None
        lab = L.kmeans_assign(x1, self.centers[n], x2).view(b, h, w)
        maps = L.onehot_nearest(lab, self.clusters_per_layer[n], self.out_size, self.out_size, out=out,
                                off=-1.0 if pm1 else 0.0)
        return lab.view(b, 1, h, w), maps

    def predict(self, features: List[torch.Tensor], channels_last_views=True):
        """features: the generator's feature list (13 maps, [B,C,H,W] channels_last views as
        returned by `Generator.forward`, or NHWC tensors).  Returns (maps [B, sum K, out, out],
        [labels per layer])."""
        nhwc = []
        for f in features:
            if channels_last_views and f.dim() == 4 and f.stride(1) == 1:
                nhwc.append(f.permute(0, 2, 3, 1))
            else:
                nhwc.append(f.permute(0, 2, 3, 1).contiguous() if channels_last_views else f)
        labs = []
        b = nhwc[1].shape[0]
        maps = torch.empty((b, sum(self.clusters_per_layer), self.out_size, self.out_size), dtype=torch.float32,
                           device=self.device)
        off = 0
        for n in range(self.n_layers):          # every layer writes its K channels straight into the concatenation
            pair = [nhwc[2 * n + 1].contiguous().float(), nhwc[2 * n + 2].contiguous().float()]
            lab, _ = self._layerwise_predict(pair, n, out=maps[:, off:off + self.clusters_per_layer[n]])
            off += self.clusters_per_layer[n]
            labs.append(lab)
        return maps, labs
