from .swav_clustering import SwAVClustering  # noqa: F401
from .one_shot_segmentor import OneShotSegmentor  # noqa: F401
