from .simclr_clustering import SimCLRClustering, SimCLRHead, simclr_train_step  # noqa: F401
