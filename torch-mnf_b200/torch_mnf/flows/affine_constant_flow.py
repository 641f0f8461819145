"""Per-dimension scale + shift flows (reference: flows/affine_constant_flow.py:7-50)."""

from __future__ import annotations

import torch
from torch import nn

from .. import _lib
from .._program import new_op
from ._base import Flow


class AffineConstantFlow(Flow):
    """x = z * exp(s) + t with learned per-dimension constants; ``scale=False`` /
    ``shift=False`` freeze the respective constant at zero (NICE's scaling layer)."""

    def __init__(self, dim: int, scale: bool = True, shift: bool = True) -> None:
        super().__init__()
        self.dim = dim
        for name, learned in (("s", scale), ("t", shift)):
            if learned:
                setattr(self, name, nn.Parameter(torch.randn(1, dim)))
            else:  # fixed zeros, absent from the state_dict like the reference's plain tensor
                self.register_buffer(name, torch.zeros(1, dim), persistent=False)

    def _emit(self, pk):
        return new_op(_lib.OP_AFFINE_CONST, aux_off=pk.add(self.s, self.t))

    def _shape_log_det(self, ld):
        return ld[:1]  # the reference returns sum(s, dim=1) of a [1, dim] tensor -> shape [1]


class ActNormFlow(AffineConstantFlow):
    """AffineConstantFlow with data-dependent initialisation on the first ``inverse`` call:
    s = log std(x), t = mean(x * exp(s)) (affine_constant_flow.py:42-50)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.data_dep_init_done = False

    def _init_pending(self, inverse: bool) -> bool:
        return inverse and not self.data_dep_init_done

    def _before_run(self, x, inverse):
        if not self._init_pending(inverse):
            return
        x = _lib.require_cuda_f32(x, "input")
        do_s = bool((self.s != 0).any())  # an all-zero parameter is left alone (:45,47)
        do_t = bool((self.t != 0).any())
        if do_s or do_t:
            ws = torch.empty(4 * self.dim, dtype=torch.float64, device=x.device)
            with torch.no_grad(), torch.cuda.device(x.device):
                rc = _lib.lib().mnf_actnorm_init(
                    x.data_ptr(), x.size(0), self.dim, self.s.data_ptr(), self.t.data_ptr(),
                    int(do_s), int(do_t), ws.data_ptr(), _lib.stream_ptr(x.device),
                )
            _lib.check(rc, "mnf_actnorm_init")
            self.__dict__["_program_salt"] = self.__dict__.get("_program_salt", 0) + 1
        self.data_dep_init_done = True

    def inverse(self, x):
        self._before_run(x, True)
        return super().inverse(x)
