"""invertiblenetworks.jl_b200 - B200 (sm_100a) implementation of the Glow training hot path of
slimgroup/InvertibleNetworks.jl behind the reference's operator interface.

Import as `import inb200` (the root-level loader maps that name onto this directory, whose name
contains a dot).  Compute lives in csrc/ -> libinb200.so (C ABI: include/inb200.h); lib.py binds
it with ctypes; glow.py mirrors the reference's Julia API on torch CUDA tensors.
"""
from . import lib
from . import dp
from .lib import InbError, PRECISIONS
from .glow import (ADAM, ActNorm, Conv1x1, CouplingLayerGlow, NetworkConditionalGlow, NetworkGlow, NetworkGlow3D,
                   Parameter, ResidualBlock, clear_grad, get_grads, get_params, load_params, nll_grad, save_params,
                   set_params, squeeze,
                   unsqueeze)

from .hint import (CouplingLayerBasic, CouplingLayerHINT, Haar_squeeze, NetworkMultiScaleHINT, get_depth, invHaar_unsqueeze,
                   wavelet_squeeze, wavelet_unsqueeze)

ConditionalLayerGlow = CouplingLayerGlow  # same class with n_cond > 0 (conditional_layer_glow.jl:61-66)

__all__ = [
    "ADAM", "ActNorm", "Conv1x1", "CouplingLayerGlow", "ConditionalLayerGlow", "NetworkConditionalGlow", "NetworkGlow",
    "NetworkGlow3D", "Parameter", "ResidualBlock", "clear_grad", "get_grads", "get_params", "nll_grad",
    "set_params", "save_params", "load_params", "squeeze", "unsqueeze", "CouplingLayerBasic", "CouplingLayerHINT", "NetworkMultiScaleHINT", "Haar_squeeze",
    "invHaar_unsqueeze", "wavelet_squeeze", "wavelet_unsqueeze", "get_depth", "InbError", "PRECISIONS", "lib", "dp",
]
