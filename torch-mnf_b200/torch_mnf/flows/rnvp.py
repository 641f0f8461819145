"""MNF-paper RNVP with a random Bernoulli mask per element and per call
(reference: flows/rnvp.py:7-39).  Forward only -- the reference has no inverse."""

from torch import nn

from ..models.mlp import MLP
from ._base import Flow


class RNVP(Flow):
    """mask ~ Bernoulli(0.5); y = net(mask*z); gate = sigmoid(s(y));
    x = (1-mask)*z*gate + (1-gate)*t(y) + mask*z;  log_det = sum (1-mask)*log gate."""

    def __init__(self, dim, h_sizes=(30,)):
        super().__init__()
        self.dim = dim
        self.net = MLP(dim, *h_sizes)
        self.t = nn.Linear(h_sizes[-1], dim)
        self.s = nn.Linear(h_sizes[-1], dim)

    def forward(self, z, noise=None):
        xs, ld = rnvp_stack_forward([self], z, noise, False)
        return xs[-1], ld

    def inverse(self, x):
        raise NotImplementedError("RNVP has no inverse (reference: flows/rnvp.py)")


def rnvp_stack_forward(flows, z, noise, want_inter):
    from ..layers._mnf_ops import rnvp_stack

    return rnvp_stack(flows, z, noise, want_inter)
