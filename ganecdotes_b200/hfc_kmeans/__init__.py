"""k-means baseline (baseline/hfc_kmeans of the reference): assignment / fit kernels and the preprocessor drop-in;
`preprocessor` mirrors the plugin socket `baseline/hfc_kmeans/base.py`."""
from .hfc_kmeans_clustering import FlatKMeansAssign, kmeans_fit  # noqa: F401
from .segmentor import HFCPreprocessor  # noqa: F401

preprocessor = HFCPreprocessor
