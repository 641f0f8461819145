"""Models under the reference's public names (models/__init__.py:1-4).

MLP is the conditioner container of the coupling / spline flows (its Linear layers are packed into the flow program);
MNFLeNet and MNFFeedForward compose the MNF layers -- MNFLeNet.forward additionally drives the fused Monte-Carlo
pipeline (conv moments once per image, implicit-GEMM conv2, tensor-core fc1).  ``LeNet``, the deterministic comparison
network of the reference's README plot, is outside the hot path and not provided."""

from .mlp import MLP
from .mnf_feed_forward import MNFFeedForward
from .mnf_lenet import MNFLeNet

__all__ = ["MLP", "MNFFeedForward", "MNFLeNet"]
