"""Toy 2-D densities used by the reference's tests and notebooks (reference: torch_mnf/data.py:21-31).
Host-side helpers; the samples are CPU tensors, move them to the GPU before calling a flow."""

import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # reference: torch_mnf/utils.py:13, re-exported by data


def sample_moons(n_samples: int) -> torch.Tensor:
    """Two interleaved half-circles with N(0, 0.05) jitter, seeded like the reference (random_state=0)."""
    from sklearn.datasets import make_moons

    xy, _labels = make_moons(n_samples, noise=0.05, random_state=0)
    return torch.from_numpy(xy).to(torch.float32)


def sample_blobs(n_samples: int) -> torch.Tensor:
    """Three unit-variance Gaussians centred on the diagonal at -3, 0 and 3 (random_state=0)."""
    from sklearn.datasets import make_blobs

    xy, _labels = make_blobs(n_samples, centers=[(3, 3), (0, 0), (-3, -3)], random_state=0)
    return torch.from_numpy(xy).to(torch.float32)


def sample_siggraph(n_samples: int) -> torch.Tensor:
    raise NotImplementedError("the SIGGRAPH point cloud (data/siggraph.pkl) is not shipped with this package")
