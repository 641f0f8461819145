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

    def forward(self, z):
        x, ld, _, _ = self._single().run(z, inverse=False)
        return x, self._shape_log_det(ld)

    def inverse(self, x):
        z, ld, _, _ = self._single().run(x, inverse=True)
        return z, self._shape_log_det(ld)
