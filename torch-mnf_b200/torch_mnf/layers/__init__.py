"""MNF layers and MADE under the reference's public names (layers/__init__.py:1-3).

MNFLinear / MNFConv2d run their forward and kl_div in the CUDA kernels of libmnf_b200.so (csrc/mnf_layers.cu,
csrc/tc_gemm.cu, csrc/mnf_kl.cu; training path: csrc/mnf_train.cu through layers/_train.py).  MADE / MaskedLinear are
parameter and mask containers: the masked products themselves run inside the MAF / IAF flow kernels
(csrc/made_fast.cu, csrc/flow_generic.cu).  Host-side helpers: _mnf_ops (ctypes marshalling, noise bookkeeping),
_train (autograd Functions)."""

from .made import MADE, MaskedLinear
from .mnf_conv import MNFConv2d
from .mnf_linear import MNFLinear

__all__ = ["MADE", "MaskedLinear", "MNFConv2d", "MNFLinear"]
