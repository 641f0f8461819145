"""The claim behind csrc/flow_pl.cu, checked on the CPU: a conditioner MLP of one scalar IS its piecewise-linear table.
tests/pl_reference.py restates the builder's algorithm in numpy; here its tables are compared with the layer-by-layer
evaluation on the reference's recorded weights (tests/golden) and on seeded random nets, at points inside, at and far
beyond the breakpoints.  (The CUDA builder is compared with the same restatement in tests/test_flow_pl_gpu.py.)"""

import numpy as np
import pytest
import torch

from tests import pl_reference as plr
from tests.helpers import golden_sd, golden_spec, load_golden


def _check_group(nets):
    bp = plr.breakpoints(nets)
    tab = plr.tables(nets, bp)
    rng = np.random.default_rng(0)
    span = 1.0 + (np.abs(bp).max() if len(bp) else 1.0)
    c = np.concatenate([rng.uniform(-span, span, 4000), rng.normal(size=2000), bp, bp + 1e-9, bp - 1e-9, [-1e6, 1e6, 0.0]])
    piece = np.searchsorted(bp, c, side="left")  # number of breakpoints < c: the kernel's `c > bp` count
    for q, (Ws, bs) in enumerate(nets):
        ref = plr.mlp(Ws, bs, c)
        A = np.stack([tab[i][q][0] for i in piece], axis=1)
        B = np.stack([tab[i][q][1] for i in piece], axis=1)
        got = A * c[None, :] + B
        scale = np.abs(A * c[None, :]).max() + np.abs(B).max() + 1.0
        assert np.abs(got - ref).max() <= 1e-11 * scale, (np.abs(got - ref).max(), len(bp))
    return len(bp)


@pytest.mark.parametrize("name", ["rnvp9_moons", "nsfcl3_stack"])
def test_tables_reproduce_golden_conditioners(name):
    g = load_golden(name)
    sd, specs = golden_sd(g), golden_spec(g)
    counts = []
    for i, spec in enumerate(specs):
        if spec["type"] in ("NSF_CL", "AffineHalfFlow"):
            for nets in plr.nets_of_flow(sd, i, spec):
                counts.append(_check_group(nets))
    assert counts and max(counts) <= 511, counts  # the capacity of a table in flow_pl.cu


@pytest.mark.parametrize("sizes", [(1, 16, 16, 16, 23), (1, 24, 24, 24, 1), (1, 8, 14), (1, 40, 1), (1, 32, 16, 1), (1, 12, 12, 12, 12, 12, 3)])
@pytest.mark.parametrize("scale", [0.3, 2.0])
def test_tables_reproduce_random_mlps(sizes, scale):
    g = torch.Generator().manual_seed(hash(sizes) % 1000)
    Ws = [(scale * torch.randn(o, i, generator=g, dtype=torch.float64) / max(1.0, i ** 0.5)).numpy() for i, o in zip(sizes[:-1], sizes[1:])]
    bs = [(scale * torch.randn(o, generator=g, dtype=torch.float64)).numpy() for o in sizes[1:]]
    n = _check_group([(Ws, bs)])
    assert n >= sizes[1] - 1  # every first-layer unit with a non-zero weight contributes its kink


def test_degenerate_nets():
    """Zero first-layer weights (no kink from that unit), an all-zero net, duplicated units (coincident kinks)."""
    W0 = np.array([[0.0], [1.0], [1.0], [-2.0]])
    b0 = np.array([0.5, -0.3, -0.3, 0.1])
    W1 = np.array([[1.0, -1.0, 0.5, 2.0], [0.0, 0.0, 0.0, 0.0]])
    b1 = np.array([0.2, -0.1])
    W2 = np.array([[1.5, -0.7]])
    b2 = np.array([0.0])
    _check_group([([W0, W1, W2], [b0, b1, b2])])
    Z = [np.zeros((3, 1)), np.zeros((2, 3))]
    assert _check_group([(Z, [np.zeros(3), np.zeros(2)])]) == 0
