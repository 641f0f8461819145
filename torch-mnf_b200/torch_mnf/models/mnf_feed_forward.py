"""Feed-forward net of MNFLinear layers (reference: models/mnf_feed_forward.py:14-38)."""

from collections.abc import Sequence
from typing import Any

from torch import nn

from ..layers import MNFLinear


class MNFFeedForward(nn.Sequential):
    """MNFLinear -> activation -> BatchNorm1d blocks; the final activation and batch norm are dropped."""

    def __init__(self, layer_sizes: Sequence[int], activation: type[nn.Module] = nn.ReLU, **kwargs: Any) -> None:
        mods = []
        for n_in, n_out in zip(layer_sizes[:-1], layer_sizes[1:]):
            mods += [MNFLinear(n_in, n_out, **kwargs), activation(), nn.BatchNorm1d(n_out)]
        super().__init__(*mods[:-2])

    def kl_div(self):
        return sum(layer.kl_div() for layer in self if hasattr(layer, "kl_div"))
