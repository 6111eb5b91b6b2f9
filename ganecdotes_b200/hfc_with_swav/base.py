"""Plugin socket of the reference (hfc_with_swav/base.py:1-2): the pipeline picks the
(segmentor, preprocessor) pair from this module and checks
`isinstance(self.preprocessor, hfc_with_swav.preprocessor)` (src/one_shot_pipeline.py:499)."""
from .one_shot_segmentor import OneShotSegmentor as segmentor  # noqa: F401
from .swav_clustering import SwAVClustering as preprocessor  # noqa: F401
