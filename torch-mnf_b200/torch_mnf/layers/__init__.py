"""MNF layers and MADE with the reference's public names (layers/__init__.py:1-3)."""

from .made import MADE, MaskedLinear
from .mnf_conv import MNFConv2d
from .mnf_linear import MNFLinear

__all__ = ["MADE", "MaskedLinear", "MNFConv2d", "MNFLinear"]
