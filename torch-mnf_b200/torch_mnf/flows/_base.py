"""Shared plumbing of the flow modules: single-flow programs and direction helpers."""

from __future__ import annotations

from torch import nn

from .._program import FlowProgram, state_without_caches


def _bump_salt(module, *_):
    module.__dict__["_program_salt"] = module.__dict__.get("_program_salt", 0) + 1


class Flow(nn.Module):
    """A flow = one descriptor (``_emit``) of an ``mnf_flow_stack_run`` program.

    ``forward(z) -> (x, log_det)`` and ``inverse(x) -> (z, log_det)`` run a one-op program;
    inside a ``NormalizingFlow`` the container fuses all ops into a single launch instead.
    """

    def __init__(self) -> None:
        super().__init__()
        # load_state_dict(assign=True) swaps the Parameter objects without touching version counters or the old
        # tensors' storage: invalidate every cached program that packed this flow's parameters
        self.register_load_state_dict_post_hook(_bump_salt)

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): buffers (MADE masks, Glow.P) are re-created
        out = super()._apply(fn, *args, **kwargs)
        _bump_salt(self)
        return out

    __getstate__ = state_without_caches  # copy.deepcopy / pickle / torch.save(model) drop the launch caches

    # hooks of flows with a data-dependent initialisation (ActNormFlow overrides them).  Class attributes, so that the
    # container's per-call getattr() is a plain class-dict hit instead of nn.Module.__getattr__'s miss path (~1 us per
    # flow and call, a fifth of a small-batch call's host time)
    _init_pending = None
    _before_run = None

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
