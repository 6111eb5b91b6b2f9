"""ganecdotes_b200: B200-native (sm_100a) per-pixel hidden-feature clustering path of
ganecdotes (StyleGAN2 synthesis -> per-pixel feature vectors -> SwAV head), behind the
reference's Python API.  Hand-written CUDA kernels reached through a C ABI
(include/ganecdotes_b200.h); no CPU fallback."""
__version__ = "0.1.0"
