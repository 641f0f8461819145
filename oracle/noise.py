"""Noise record / replay for the oracle (test infrastructure, see oracle/__init__.py).

The reference draws its noise inline with ``torch.randn_like``, ``torch.bernoulli``
and ``torch.randn`` (rnvp.py:28, mnf_linear.py:54,60,70, mnf_conv.py:76,82,110,118).
To compare two implementations on *identical* noise the draws are recorded in
call order while the reference runs and replayed, in the same order, into the
oracle and into the CUDA path.
"""

from __future__ import annotations

import contextlib

import torch


class NoiseTape:
    """Ordered list of noise tensors; ``kind`` is 'normal' or 'bernoulli'."""

    def __init__(self, draws=None):
        self.draws: list[tuple[str, torch.Tensor]] = list(draws or [])
        self.pos = 0

    # -- replay side -------------------------------------------------------
    def _next(self, kind: str, shape) -> torch.Tensor:
        if self.pos >= len(self.draws):
            raise IndexError(f"noise tape exhausted at draw {self.pos} ({kind}{tuple(shape)})")
        k, t = self.draws[self.pos]
        if k != kind or tuple(t.shape) != tuple(shape):
            raise ValueError(
                f"noise tape draw {self.pos}: recorded {k}{tuple(t.shape)}, "
                f"requested {kind}{tuple(shape)}"
            )
        self.pos += 1
        return t

    def normal(self, shape) -> torch.Tensor:
        return self._next("normal", shape)

    def bernoulli(self, shape) -> torch.Tensor:
        return self._next("bernoulli", shape)

    def rewind(self) -> "NoiseTape":
        self.pos = 0
        return self

    # -- (de)serialisation for tests/golden/*.npz --------------------------
    def to_npz_dict(self, prefix="noise/") -> dict:
        out = {}
        for i, (k, t) in enumerate(self.draws):
            out[f"{prefix}{i:03d}_{k}"] = t.detach().cpu().numpy()
        return out

    @classmethod
    def from_npz(cls, npz, prefix="noise/") -> "NoiseTape":
        keys = sorted(k for k in npz.keys() if k.startswith(prefix))
        draws = []
        for k in keys:
            kind = k.rsplit("_", 1)[1]
            draws.append((kind, torch.from_numpy(npz[k]).clone()))
        return cls(draws)


class FreshNoise:
    """Same interface as NoiseTape but draws from torch's global RNG (used when the
    oracle is timed as the CPU baseline: the RNG cost is part of the reference path)."""

    def normal(self, shape) -> torch.Tensor:
        return torch.randn(tuple(shape))

    def bernoulli(self, shape) -> torch.Tensor:
        return torch.bernoulli(torch.full(tuple(shape), 0.5))


@contextlib.contextmanager
def record(tape: NoiseTape):
    """Monkeypatch torch's RNG entry points the reference uses; append every draw to tape."""
    o_randn_like, o_bern, o_randn = torch.randn_like, torch.bernoulli, torch.randn

    def randn_like(t, *a, **k):
        r = o_randn_like(t, *a, **k)
        tape.draws.append(("normal", r.detach().clone()))
        return r

    def bernoulli(t, *a, **k):
        r = o_bern(t, *a, **k)
        tape.draws.append(("bernoulli", r.detach().clone()))
        return r

    def randn(*a, **k):
        r = o_randn(*a, **k)
        tape.draws.append(("normal", r.detach().clone()))
        return r

    torch.randn_like, torch.bernoulli, torch.randn = randn_like, bernoulli, randn
    try:
        yield tape
    finally:
        torch.randn_like, torch.bernoulli, torch.randn = o_randn_like, o_bern, o_randn
