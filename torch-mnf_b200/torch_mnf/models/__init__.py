"""Models with the reference's public names (models/__init__.py:1-4).  ``LeNet`` (the deterministic
comparison network of the reference's README plot) is outside the hot path and not provided."""

from .mlp import MLP
from .mnf_feed_forward import MNFFeedForward
from .mnf_lenet import MNFLeNet

__all__ = ["MLP", "MNFFeedForward", "MNFLeNet"]
