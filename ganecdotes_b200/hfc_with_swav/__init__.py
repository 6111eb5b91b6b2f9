from .swav_clustering import SwAVClustering  # noqa: F401
