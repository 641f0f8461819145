"""RealNVP / NICE coupling flow (reference: flows/affine_half_flow.py:20-66)."""

from __future__ import annotations

from collections.abc import Sequence

from .. import _lib
from .._program import check_leaky, net_tensors, new_op
from ..models.mlp import MLP
from ._base import Flow


class AffineHalfFlow(Flow):
    """Half of the dimensions are scaled by exp(s) and shifted by t, both MLPs of the other
    half; ``parity`` picks which half conditions.  ``scale=False`` gives NICE (shift only)."""

    def __init__(self, dim: int, parity: bool, h_sizes: Sequence[int] = (24, 24, 24), scale: bool = True,
                 shift: bool = True) -> None:
        super().__init__()
        if dim % 2:
            raise ValueError("AffineHalfFlow needs an even dim")
        self.dim, self.parity = dim, parity
        self.s_net = MLP(dim // 2, *h_sizes, dim // 2) if scale else None
        self.t_net = MLP(dim // 2, *h_sizes, dim // 2) if shift else None
        self._sizes = (dim // 2, *h_sizes, dim // 2)

    def _emit(self, pk):
        check_leaky(self.s_net, self.t_net)
        flags = _lib.FLAG_PARITY if self.parity else 0
        offs = [0, 0]
        for i, (net, flag) in enumerate(((self.s_net, _lib.FLAG_SCALE), (self.t_net, _lib.FLAG_SHIFT))):
            if net is not None:
                flags |= flag
                offs[i] = pk.add(*net_tensors(net.linears()))
        return new_op(_lib.OP_AFFINE_HALF, flags=flags, sizes=self._sizes, net_off=offs)

    def forward(self, z, inverse: bool = False):
        if inverse:
            return self.inverse(z)
        return super().forward(z)
