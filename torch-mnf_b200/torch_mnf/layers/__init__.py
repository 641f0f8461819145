"""MNF layers and MADE with the reference's public names (layers/__init__.py:1-3)."""

from .made import MADE, MaskedLinear

__all__ = ["MADE", "MaskedLinear"]
