"""Masked autoregressive flows (reference: flows/maf.py:21-72)."""

from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import nn

from .. import _lib
from .._program import net_tensors, new_op
from ._base import Flow


class MAF(Flow):
    """``inverse`` (density) is one MADE pass: z = x*exp(s)+t, dims flipped afterwards if
    ``parity``; ``forward`` (sampling) decodes the D dimensions sequentially."""

    _sequential_on_forward = True
    #: "auto": TF32 tensor cores for large batches (2e-3 tolerance class), exact fp32 otherwise; "fp32"; "tf32"
    precision = "auto"

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

    def _density(self, v):
        """One-pass direction (MAF.inverse / IAF.forward)."""
        if _use_tc([self], v):
            from ..layers.made import MadeStackPlan, made_density

            plan = self.__dict__.setdefault("_tc_plan", MadeStackPlan([self]))
            z, ld, _ = made_density(plan, v)
            return z, ld
        return None

    def inverse(self, x):
        out = self._density(x) if self._sequential_on_forward else None
        return out if out is not None else super().inverse(x)

    def forward(self, z):
        out = self._density(z) if not self._sequential_on_forward else None
        return out if out is not None else super().forward(z)


class IAF(MAF):
    """MAF with the two directions swapped: fast sampling, D-pass density (maf.py:65-72)."""

    _sequential_on_forward = False


def _use_tc(flows, v) -> bool:
    """Tensor-core density path?  All flows must be eligible MAFs with the same precision policy."""
    from ..layers.made import TC_MIN_ELEMS, made_tc_eligible

    if not flows or not all(isinstance(f, MAF) and made_tc_eligible(f) for f in flows):
        return False
    if torch.is_grad_enabled() and (v.requires_grad or any(p.requires_grad for f in flows for p in f.parameters())):
        return False  # training goes through the differentiable program (mnf_flow_stack_backward)
    prec = {f.precision for f in flows}
    if prec == {"fp32"}:
        return False
    if prec == {"tf32"}:
        return True
    if all(f.dim == 64 and list(getattr(f.net, "hidden_sizes", [])) == [24, 24, 24] for f in flows):
        return False  # the exact-fp32 constant-bank MADE kernel (made_fast.cu) beats the TF32 GEMM chain here
    return "fp32" not in prec and v.is_cuda and v.numel() >= TC_MIN_ELEMS
