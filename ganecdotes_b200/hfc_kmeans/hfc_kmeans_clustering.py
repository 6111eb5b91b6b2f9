"""Per-pixel k-means assignment on GAN features (BASELINE config 5) - the `predict` side of the
reference's `FlatKMeansHFC` (baseline/hfc_kmeans/hfc_kmeans_clustering.py:169-230).

For each of the `n_layers` resolutions, the two same-resolution feature maps f[2n+1], f[2n+2]
(ref lib/oneshot/image_augmentor.py:80-90 with skip_const=True) form the per-pixel vector; every
pixel is assigned to its nearest cluster centre (first index on ties, like sklearn's argmin);
labels become one-hot maps resized NEAREST to `out_size` and concatenated over layers.
`fit` (Lloyd / k-means++ inside scikit-learn) stays with scikit-learn: pass the fitted
`cluster_centers_` (SURVEY.md §8(f) lists a native fit as "next").
"""
from typing import List, Sequence

import torch

from .. import _lib as L


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

    def _layerwise_predict(self, feats_nhwc: List[torch.Tensor], n: int):
        """ref :169-208.  feats_nhwc: the map(s) of layer n as fp32 NHWC tensors.
        Returns (labels int32 [b,1,h,w], label_maps float [b,K,out,out])."""
        f1 = feats_nhwc[0]
        b, h, w, c1 = f1.shape
        x1 = f1.reshape(-1, c1)
        x2 = feats_nhwc[1].reshape(-1, feats_nhwc[1].shape[3]) if len(feats_nhwc) > 1 else None
        lab = L.kmeans_assign(x1, self.centers[n], x2).view(b, h, w)
        maps = L.onehot_nearest(lab, self.clusters_per_layer[n], self.out_size, self.out_size)
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
        outs, labs = [], []
        for n in range(self.n_layers):
            pair = [nhwc[2 * n + 1].contiguous().float(), nhwc[2 * n + 2].contiguous().float()]
            lab, maps = self._layerwise_predict(pair, n)
            outs.append(maps)
            labs.append(lab)
        return torch.cat(outs, 1), labs
