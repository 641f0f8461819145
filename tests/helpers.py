"""Shared helpers for the parity tests (golden loading, spec handling)."""

import json
import os

import numpy as np
import torch

from oracle.noise import NoiseTape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_sd(npz, prefix="sd/"):
    return {k[len(prefix):]: torch.from_numpy(npz[k]).clone() for k in npz.keys() if k.startswith(prefix)}


def golden_spec(npz):
    return json.loads(str(npz["spec"]))


def golden_tape(npz, prefix):
    return NoiseTape.from_npz(npz, prefix)


def t(npz, key):
    return torch.from_numpy(npz[key]).clone()


def max_rel(a, b, atol=0.0):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / (b.abs() + atol + 1e-30)).max())
