"""Shared plumbing of the flow modules: single-flow programs and direction helpers."""

from __future__ import annotations

from torch import nn

from .._program import FlowProgram


class Flow(nn.Module):
    """A flow = one descriptor (``_emit``) of an ``mnf_flow_stack_run`` program.

    ``forward(z) -> (x, log_det)`` and ``inverse(x) -> (z, log_det)`` run a one-op program;
    inside a ``NormalizingFlow`` the container fuses all ops into a single launch instead.
    """

    def _emit(self, pk):  # -> _lib.FlowOp
        raise NotImplementedError

    def _single(self) -> FlowProgram:
        prog = self.__dict__.get("_single_prog")
        if prog is None:
            prog = FlowProgram([self])
            self.__dict__["_single_prog"] = prog
        return prog

    def _shape_log_det(self, ld):
        return ld

    def _run_single(self, v, inverse: bool):
        prog = self._single()
        hook = getattr(self, "_before_run", None)
        if prog.needs_grad(v):  # training: keep the autograd graph (mnf_flow_stack_backward)
            if hook is not None:
                hook(v, inverse)
            ld, inter = prog.run_autograd(v, inverse)
            return inter[-1], self._shape_log_det(ld)
        out, ld, _, _ = prog.run(v, inverse=inverse)
        return out, self._shape_log_det(ld)

    def forward(self, z):
        return self._run_single(z, inverse=False)

    def inverse(self, x):
        return self._run_single(x, inverse=True)
