"""Invertible 1x1 'convolution' with PLU parametrisation (reference: flows/glow.py:5-37)."""

import torch
from torch import nn

from .. import _lib
from .._program import new_op
from ._base import Flow


class Glow(Flow):
    """x = z @ W with W = P (tril(L,-1)+I) (triu(U,1)+diag S); log_det = sum log|S|.

    W, W^-1 and log_det are assembled on the device by ``mnf_glow_assemble`` into the
    program's parameter blob whenever L, S or U change; the per-point D x D product runs
    inside the fused flow kernel."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        Q = nn.init.orthogonal_(torch.randn(dim, dim))
        P, L, U = torch.linalg.lu(Q)
        self.register_buffer("P", P, persistent=False)  # fixed permutation, not in the state_dict
        self.L = nn.Parameter(L)
        self.S = nn.Parameter(U.diag())
        self.U = nn.Parameter(torch.triu(U, diagonal=1))

    def _emit(self, pk):
        D = self.dim

        def fill(blob, off):
            dev = blob.device
            args = [t.detach().to(torch.float32).contiguous() for t in (self.P, self.L, self.U, self.S)]
            for t in args:
                if t.device != dev:
                    raise RuntimeError(f"Glow parameter on {t.device}, input on {dev}")
            with torch.cuda.device(dev):
                rc = _lib.lib().mnf_glow_assemble(
                    *(t.data_ptr() for t in args), blob[off:].data_ptr(), D, _lib.stream_ptr(dev)
                )
            _lib.check(rc, "mnf_glow_assemble")

        return new_op(_lib.OP_GLOW, aux_off=pk.reserve(2 * D * D + 1, fill, self._assemble_torch))

    def _assemble_torch(self):
        """[W | W^-1 | log_det] as a differentiable function of L, S, U (training path only; glow.py:20-24,34-35)."""
        eye = torch.eye(self.dim, device=self.L.device, dtype=self.L.dtype)
        W = self.P @ (torch.tril(self.L, diagonal=-1) + eye) @ (torch.triu(self.U, diagonal=1) + torch.diag(self.S))
        return torch.cat([W.reshape(-1), torch.inverse(W).reshape(-1), self.S.abs().log().sum().reshape(1)])

    def _shape_log_det(self, ld):
        return ld[0]  # the reference returns a 0-d tensor
