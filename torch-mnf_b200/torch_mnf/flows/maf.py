"""Masked autoregressive flows (reference: flows/maf.py:21-72)."""

from __future__ import annotations

from collections.abc import Sequence

from torch import nn

from .. import _lib
from .._program import net_tensors, new_op
from ._base import Flow


class MAF(Flow):
    """``inverse`` (density) is one MADE pass: z = x*exp(s)+t, dims flipped afterwards if
    ``parity``; ``forward`` (sampling) decodes the D dimensions sequentially."""

    _sequential_on_forward = True

    def __init__(self, dim: int, parity: bool, net: nn.Module | None = None,
                 h_sizes: Sequence[int] = (24, 24, 24)) -> None:
        super().__init__()
        from ..layers.made import MADE

        self.dim, self.parity = dim, parity
        self.net = net or MADE(dim, h_sizes, 2 * dim, natural_ordering=True)

    def _emit(self, pk):
        from ..layers.made import MaskedLinear

        lin = [m for m in self.net if isinstance(m, MaskedLinear)]
        sizes = [lin[0].in_features] + [m.out_features for m in lin]
        if sizes[0] != self.dim or sizes[-1] != 2 * self.dim:
            raise ValueError("MAF needs a MADE with n_in = dim and n_out = 2*dim")
        flags = (_lib.FLAG_PARITY if self.parity else 0) | (
            _lib.FLAG_MADE_SEQ if self._sequential_on_forward else 0
        )
        off = pk.add(*net_tensors(lin, masks=[m.mask for m in lin]))
        return new_op(_lib.OP_MADE, flags=flags, sizes=sizes, net_off=(off, 0))

    def inverse(self, x):
        if not self._sequential_on_forward or not _dense_made_ok(self):
            return super().inverse(x)
        from ..layers.made import made_density

        return made_density([self], x)

    def forward(self, z):
        if self._sequential_on_forward or not _dense_made_ok(self):
            return super().forward(z)
        from ..layers.made import made_density

        return made_density([self], z)


class IAF(MAF):
    """MAF with the two directions swapped: fast sampling, D-pass density (maf.py:65-72)."""

    _sequential_on_forward = False


def _dense_made_ok(flow) -> bool:
    return False  # the tiled MADE kernel registers itself here (layers/made.py)
