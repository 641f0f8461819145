"""Masked autoregressive MLP (reference: layers/made.py:11-94; Germain et al. 2015)."""

from __future__ import annotations

import numpy as np
import torch
from torch import nn


class MaskedLinear(nn.Linear):
    """Dense layer whose weight is multiplied by a fixed 0/1 connectivity mask ``[n_in, n_out]``.
    Inside MAF/IAF the mask is folded into the weight when the flow program is packed."""

    def __init__(self, n_in, n_out, bias=True):
        super().__init__(n_in, n_out, bias)
        self.register_buffer("mask", torch.ones(n_in, n_out, dtype=torch.bool))

    def set_mask(self, mask: np.ndarray) -> None:
        # in place: bumps the buffer's version so cached flow programs re-fold the mask
        self.mask.copy_(torch.from_numpy(np.ascontiguousarray(mask)).to(self.mask.device, torch.bool))

    def forward(self, x):
        raise RuntimeError(
            "MaskedLinear is a parameter container in the B200 build; MADE runs inside the "
            "fused MAF/IAF kernels (flows.MAF / flows.IAF)"
        )


class MADE(nn.Sequential):
    """MaskedLinear / ReLU chain with autoregressive masks.

    Degrees m(k) follow the reference: inputs keep their natural order (or a permutation),
    hidden unit degrees are drawn with numpy's legacy ``RandomState(seed)`` from
    [min degree of previous layer, n_in - 1), hidden masks use ``<=``, the output mask ``<``
    and is tiled when ``n_out`` is a multiple of ``n_in`` (made.py:59-94)."""

    def __init__(self, n_in, hidden_sizes, n_out, num_masks=1, natural_ordering=False):
        if n_out % n_in:
            raise AssertionError("n_out must be integer multiple of n_in")
        widths = [n_in, *hidden_sizes, n_out]
        mods = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [MaskedLinear(a, b), nn.ReLU()]
        super().__init__(*mods[:-1])
        self.n_in, self.n_out, self.hidden_sizes = n_in, n_out, list(hidden_sizes)
        self.natural_ordering, self.num_masks = natural_ordering, num_masks
        self.seed = 0
        self.m = {}
        self.update_masks()

    def update_masks(self):
        if self.m and self.num_masks == 1:
            return
        rng = np.random.RandomState(self.seed)
        self.seed = (self.seed + 1) % self.num_masks
        degrees = {-1: np.arange(self.n_in) if self.natural_ordering else rng.permutation(self.n_in)}
        for i, width in enumerate(self.hidden_sizes):
            degrees[i] = rng.randint(degrees[i - 1].min(), self.n_in - 1, size=width)
        self.m = degrees
        n_hidden = len(self.hidden_sizes)
        masks = [degrees[i - 1][:, None] <= degrees[i][None, :] for i in range(n_hidden)]
        out_mask = degrees[n_hidden - 1][:, None] < degrees[-1][None, :]
        masks.append(np.tile(out_mask, (1, self.n_out // self.n_in)))
        for layer, mask in zip((m for m in self if isinstance(m, MaskedLinear)), masks):
            layer.set_mask(mask)


def made_density(flows, x):
    raise NotImplementedError
