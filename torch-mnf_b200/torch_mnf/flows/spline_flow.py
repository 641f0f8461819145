"""Neural spline flows, coupling (NSF_CL) and autoregressive (NSF_AR)
(reference: flows/spline_flow.py:182-285; Durkan et al. 2019)."""

import torch
from torch import nn

from .. import _lib
from .._program import check_leaky, net_tensors, new_op, spline_edge_derivative
from ..models.mlp import MLP
from ._base import Flow


class NSF_CL(Flow):
    """Coupling layer: f1(lower) parameterises a K-bin rational-quadratic spline applied to
    upper, then f2(upper_new) one applied to lower; tails outside [-B, B] are the identity."""

    def __init__(self, dim, K=5, B=3, n_h=8, net_class=MLP):
        super().__init__()
        if dim % 2:
            raise ValueError("NSF_CL needs an even dim")
        self.dim, self.K, self.B = dim, K, B
        n_out = (3 * K - 1) * dim // 2
        self.f1 = net_class(dim // 2, n_h, n_h, n_h, n_out)
        self.f2 = net_class(dim // 2, n_h, n_h, n_h, n_out)

    def _emit(self, pk):
        check_leaky(self.f1, self.f2)
        lin1, lin2 = self.f1.linears(), self.f2.linears()
        sizes = [lin1[0].in_features] + [m.out_features for m in lin1]
        offs = (pk.add(*net_tensors(lin1)), pk.add(*net_tensors(lin2)))
        return new_op(_lib.OP_NSF_CL, K=self.K, bound=self.B, sizes=sizes, net_off=offs,
                      edge_deriv=spline_edge_derivative())


class NSF_AR(Flow):
    """Autoregressive layer: dim i is transformed by a spline whose parameters come from an
    MLP of dims < i (a learned constant for dim 0).  ``forward`` runs the spline inverse and
    conditions on its own outputs, ``inverse`` the spline forward (spline_flow.py:199-235)."""

    def __init__(self, dim, K=5, B=3, n_h=8, net_class=MLP):
        super().__init__()
        self.dim, self.K, self.B = dim, K, B
        self.layers = nn.ModuleList()
        self.init_param = nn.Parameter(torch.Tensor(3 * K - 1))
        for i in range(1, dim):
            self.layers += [net_class(i, n_h, n_h, n_h, 3 * K - 1)]
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.uniform_(self.init_param, -1 / 2, 1 / 2)

    def _emit(self, pk):
        check_leaky(*self.layers)
        aux = pk.add(self.init_param)
        sizes, off = [1, 3 * self.K - 1], 0  # dim == 1: no conditioner, descriptor unused
        if len(self.layers):
            lin0 = self.layers[0].linears()
            sizes = [1] + [m.out_features for m in lin0]
            tensors = []
            for net in self.layers:
                tensors += net_tensors(net.linears())
            off = pk.add(*tensors)  # nets back to back: the kernel walks them by size
        return new_op(_lib.OP_NSF_AR, K=self.K, bound=self.B, sizes=sizes, net_off=(off, 0), aux_off=aux,
                      edge_deriv=spline_edge_derivative())
