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
        got = density_stack([self], v, self.__dict__, want_inter=False)
        return None if got is None else (got[0], got[1])

    def inverse(self, x):
        out = self._density(x) if self._sequential_on_forward else None
        return out if out is not None else super().inverse(x)

    def forward(self, z):
        out = self._density(z) if not self._sequential_on_forward else None
        return out if out is not None else super().forward(z)


class IAF(MAF):
    """MAF with the two directions swapped: fast sampling, D-pass density (maf.py:65-72)."""

    _sequential_on_forward = False


def _tc_mode(flows, v):
    """Which density path a stack of MAF flows takes in its one-pass direction: None = the exact-fp32 flow program,
    "fused" = the single persistent tcgen05 kernel (dim 64, hidden <= 31: csrc/made_fused.cu), "chain" = one TF32 GEMM
    per layer (any other eligible shape).  Both tensor-core paths are in the 2e-3 tolerance class BASELINE.json states
    for the MADE GEMMs; ``precision = "fp32"`` on the flows keeps the exact path."""
    from ..layers.made import FUSED_MAX_FLOWS, TC_MIN_ELEMS, made_fused_eligible, made_tc_eligible

    if not flows or not all(isinstance(f, MAF) and made_tc_eligible(f) for f in flows):
        return None
    if torch.is_grad_enabled() and (v.requires_grad or any(p.requires_grad for f in flows for p in f.parameters())):
        return None  # training goes through the differentiable program (mnf_flow_stack_backward)
    prec = {f.precision for f in flows}
    if "fp32" in prec or not v.is_cuda:
        return None
    fused = (len(flows) <= FUSED_MAX_FLOWS and all(made_fused_eligible(f) for f in flows)
             and len({len(f.net.hidden_sizes) for f in flows}) == 1)
    if prec == {"tf32"}:
        return "fused" if fused else "chain"
    if v.numel() < TC_MIN_ELEMS:
        return None
    return "fused" if fused else "chain"


def _use_tc(flows, v) -> bool:
    return _tc_mode(flows, v) is not None


def density_stack(flows, v, cache, want_inter, want_z=True, want_log_prob=False, log_prob_out=None):
    """Runs `flows` (execution order) in their one-pass direction on a tensor-core path if one applies.
    -> (z or None, log_det, intermediates or None, log_prob or None), or None when the exact program should run.
    `cache`: the owning module's __dict__ (packed plans are kept there, dropped by __getstate__)."""
    mode = _tc_mode(flows, v)
    if mode is None:
        return None
    from ..layers import made as M

    entry = cache.get("_tc_plan")
    ids = tuple(id(f) for f in flows)
    if entry is None or entry[0] != (mode, ids):
        plans = ((M.FusedMadePlan(flows, True), M.FusedMadePlan(flows, False)) if mode == "fused"
                 else M.MadeStackPlan(flows))
        entry = ((mode, ids), plans)
        cache["_tc_plan"] = entry
    if mode == "fused":
        return M.made_density_fused(entry[1], v, want_inter=want_inter, want_z=want_z, want_log_prob=want_log_prob,
                                    log_prob_out=log_prob_out)
    z, ld, inter = M.made_density(entry[1], v, want_inter=want_inter)
    lp = None
    if want_log_prob:
        lp = ld - 0.5 * z.square().sum(1) - 0.5 * z.size(1) * 1.8378770664093453
    return z, ld, inter, lp
