"""Plugin socket of the reference (hfc_with_swav/base.py:1-2): the pipeline picks the
(segmentor, preprocessor) pair from this module and checks
`isinstance(self.preprocessor, hfc_with_swav.preprocessor)` (src/one_shot_pipeline.py:499)."""
from .swav_clustering import SwAVClustering as preprocessor  # noqa: F401

try:  # the one-shot segmentor head is a "next" row; re-exported when available
    from .segmentor import OneShotSegmentor as segmentor  # noqa: F401
except ImportError:  # pragma: no cover
    segmentor = None
