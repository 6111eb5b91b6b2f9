"""Drop-in for the reference's `models/stylegan2/op` package (and for the
`lib/gan/optim` autograd wrappers' forward): same names, same signatures."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu, fused_bias_act
from .upfirdn2d import upfirdn2d, upfirdn2d_native_layout
from . import conv2d_gradfix  # noqa: F401  (lib/gan/optim/conv2d_gradfix.py: conv2d, conv_transpose2d, no_weight_gradients)

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "fused_bias_act", "upfirdn2d", "upfirdn2d_native_layout", "conv2d_gradfix"]
