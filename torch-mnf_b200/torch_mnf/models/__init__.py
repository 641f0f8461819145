"""Models with the reference's public names (models/__init__.py:1-4)."""

from .mlp import MLP

__all__ = ["MLP"]
