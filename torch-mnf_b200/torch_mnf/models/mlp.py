"""Conditioner MLP (parameter container with the reference's layout, models/mlp.py:4-12)."""

from torch import nn


class MLP(nn.Sequential):
    """Linear / LeakyReLU(leaky_a) chain without a trailing activation.

    State-dict keys are ``{2*i}.weight`` / ``{2*i}.bias`` as in the reference.  Inside the
    flows the network is evaluated by the fused CUDA kernels, which read these parameters.
    """

    def __init__(self, *layer_sizes, leaky_a=0.2):
        mods = []
        for n_in, n_out in zip(layer_sizes[:-1], layer_sizes[1:]):
            mods += [nn.Linear(n_in, n_out), nn.LeakyReLU(leaky_a)]
        super().__init__(*mods[:-1])
        self.layer_sizes = tuple(layer_sizes)
        self.leaky_a = leaky_a

    def linears(self):
        return [m for m in self if isinstance(m, nn.Linear)]
