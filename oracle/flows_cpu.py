"""CPU oracle: flow forward / inverse with log-det accumulation.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional restatement of the
reference's flow modules on PyTorch-CPU fp32.  Every function takes a plain
``dict`` of tensors whose keys are the reference's ``state_dict`` names with the
module prefix stripped, plus a small JSON-able ``spec`` dict
(``{"type": "NSF_CL", "dim": 2, "K": 8, "B": 3}``), and returns ``(out, log_det)``
exactly like the reference module method it restates.

Citations are ``file:line`` in the reference checkout (torch_mnf/...).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

MIN_W = 1e-3  # spline_flow.py:17
MIN_H = 1e-3  # spline_flow.py:18
MIN_D = 1e-3  # spline_flow.py:19


def sub(sd: dict, prefix: str) -> dict:
    """Entries of ``sd`` under ``prefix`` with the prefix removed."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ---------------------------------------------------------------------------
# conditioners
# ---------------------------------------------------------------------------
def _layer_ids(p: dict) -> list[int]:
    return sorted({int(k.split(".")[0]) for k in p if k.split(".")[0].isdigit()})


def mlp(p: dict, x: torch.Tensor, leaky: float = 0.2) -> torch.Tensor:
    """models/mlp.py:4-12 -- Linear/LeakyReLU(0.2) chain, no activation after the last Linear."""
    ids = _layer_ids(p)
    for n, i in enumerate(ids):
        x = F.linear(x, p[f"{i}.weight"], p[f"{i}.bias"])
        if n + 1 < len(ids):
            x = F.leaky_relu(x, leaky)
    return x


def made(p: dict, x: torch.Tensor) -> torch.Tensor:
    """layers/made.py:22-23 (MaskedLinear) chained with ReLU as in made.py:43-47."""
    ids = _layer_ids(p)
    for n, i in enumerate(ids):
        x = x @ (p[f"{i}.weight"].T * p[f"{i}.mask"]) + p[f"{i}.bias"]
        if n + 1 < len(ids):
            x = torch.relu(x)
    return x


def made_masks(n_in: int, hidden: list[int], n_out: int, natural: bool, seed: int = 0):
    """layers/made.py:59-94 -- connectivity masks from numpy's legacy RandomState(seed)."""
    import numpy as np

    rng = np.random.RandomState(seed)
    deg = {-1: np.arange(n_in) if natural else rng.permutation(n_in)}
    for lyr, size in enumerate(hidden):
        deg[lyr] = rng.randint(deg[lyr - 1].min(), n_in - 1, size=size)
    L = len(hidden)
    masks = [deg[lyr - 1][:, None] <= deg[lyr][None, :] for lyr in range(L)]
    last = deg[L - 1][:, None] < deg[-1][None, :]
    if n_out > n_in:
        last = np.concatenate([last] * (n_out // n_in), axis=1)
    masks.append(last)
    return masks


# ---------------------------------------------------------------------------
# affine flows
# ---------------------------------------------------------------------------
def _const_st(p: dict, spec: dict):
    d = spec["dim"]
    ref = next(iter(p.values())) if p else torch.zeros(())
    s = p["s"] if "s" in p else torch.zeros(1, d, dtype=ref.dtype)  # affine_constant_flow.py:15
    t = p["t"] if "t" in p else torch.zeros(1, d, dtype=ref.dtype)  # affine_constant_flow.py:16
    return s, t


def affine_constant(p, spec, v, inverse: bool):
    """affine_constant_flow.py:18-26."""
    s, t = _const_st(p, spec)
    if inverse:
        return (v - t) * torch.exp(-s), torch.sum(-s, dim=1)
    return v * torch.exp(s) + t, torch.sum(s, dim=1)


def actnorm_init(p, spec, x):
    """affine_constant_flow.py:44-49 -- data dependent init; returns new (s, t)."""
    s, t = _const_st(p, spec)
    if not all(s.squeeze() == 0):
        s = x.std(dim=0, keepdim=True).log()
    if not all(t.squeeze() == 0):
        t = (x * s.exp()).mean(dim=0, keepdim=True)
    return s, t


def affine_half(p, spec, v, inverse: bool):
    """affine_half_flow.py:44-62."""
    d = spec["dim"]
    v0, v1 = v.chunk(2, dim=1)
    if spec["parity"]:
        v0, v1 = v1, v0
    zeros = v0.new_zeros(v0.size(0), d // 2)  # affine_half_flow.py:38
    s = mlp(sub(p, "s_net."), v0) if spec.get("scale", True) else zeros
    t = mlp(sub(p, "t_net."), v0) if spec.get("shift", True) else zeros
    if inverse:
        o1 = (v1 - t) / s.exp()
        s = -s
    else:
        o1 = s.exp() * v1 + t
    o0 = v0
    if spec["parity"]:
        o0, o1 = o1, o0
    return torch.cat([o0, o1], dim=1), s.sum(1)


def glow_W(p):
    """glow.py:20-24."""
    n = len(p["L"])
    L = torch.tril(p["L"], diagonal=-1) + torch.eye(n, dtype=p["L"].dtype)
    U = torch.triu(p["U"], diagonal=1)
    return p["P"] @ L @ (U + p["S"].diag())


def glow(p, spec, v, inverse: bool):
    """glow.py:26-37."""
    W = glow_W(p)
    ld = p["S"].abs().log().sum()
    if inverse:
        return v @ torch.inverse(W), -ld
    return v @ W, ld


def maf_density(p, spec, x):
    """maf.py:53-62 (MAF.inverse): one MADE pass."""
    d = x.size(1)
    st = made(sub(p, "net."), x)
    s, t = st.split(d, dim=1)
    z = x * s.exp() + t
    if spec["parity"]:
        z = z.flip(dims=[1])
    return z, s.sum(1)


def maf_sample(p, spec, z):
    """maf.py:39-51 (MAF.forward): D sequential MADE passes."""
    B, d = z.shape
    x = torch.zeros_like(z)
    ld = torch.zeros(B, dtype=z.dtype)
    if spec["parity"]:
        z = z.flip(dims=[1])
    net = sub(p, "net.")
    for i in range(d):
        st = made(net, x.clone())
        s, t = st.split(d, dim=1)
        x[:, i] = (z[:, i] - t[:, i]) * torch.exp(-s[:, i])
        ld += -s[:, i]
    return x, ld


def rnvp(p, spec, z, mask):
    """rnvp.py:25-39 with the Bernoulli mask injected (rnvp.py:28 draws it inline)."""
    z1, z2 = (1 - mask) * z, mask * z
    y = mlp(sub(p, "net."), z2)
    shift = F.linear(y, p["t.weight"], p["t.bias"])
    scale = F.linear(y, p["s.weight"], p["s.bias"])
    gate = torch.sigmoid(scale)
    ld = ((1 - mask) * gate.log()).sum(1)
    x = (z1 * gate + (1 - gate) * shift) + z2
    return x, ld


# ---------------------------------------------------------------------------
# rational-quadratic splines
# ---------------------------------------------------------------------------
def _knots(unnorm, lo, hi, min_size):
    """spline_flow.py:95-102 / 106-113: softmax -> floor -> cumsum -> pad -> rescale -> pin."""
    K = unnorm.shape[-1]
    w = F.softmax(unnorm, dim=-1)
    w = min_size + (1 - min_size * K) * w
    cum = torch.cumsum(w, dim=-1)
    cum = F.pad(cum, pad=(1, 0), mode="constant", value=0.0)
    cum = (hi - lo) * cum + lo
    cum[..., 0] = lo
    cum[..., -1] = hi
    return cum, cum[..., 1:] - cum[..., :-1]


def _bin_index(knots, v, eps=1e-6):
    """spline_flow.py:22-24: in-place +eps on the last knot, then compare-and-count."""
    knots[..., -1] += eps
    return (v[..., None] >= knots).sum(dim=-1) - 1


def rqs(v, uw, uh, ud, inverse, lo, hi):
    """spline_flow.py:71-179 for inputs already known to lie in [lo, hi]."""
    cw, widths = _knots(uw, lo, hi, MIN_W)
    derivs = MIN_D + F.softplus(ud)  # spline_flow.py:104
    ch, heights = _knots(uh, lo, hi, MIN_H)

    idx = _bin_index(ch if inverse else cw, v)[..., None]  # spline_flow.py:115-118

    def pick(t):
        return t.gather(-1, idx)[..., 0]

    x_k, w_k, y_k = pick(cw), pick(widths), pick(ch)
    delta = heights / widths
    s_k = pick(delta)
    d_k = pick(derivs)
    d_k1 = pick(derivs[..., 1:])
    h_k = pick(heights)

    if inverse:  # spline_flow.py:133-162
        dy = v - y_k
        a = dy * (d_k + d_k1 - 2 * s_k) + h_k * (s_k - d_k)
        b = h_k * d_k - dy * (d_k + d_k1 - 2 * s_k)
        c = -s_k * dy
        disc = b.pow(2) - 4 * a * c
        root = (2 * c) / (-b - torch.sqrt(disc))
        out = root * w_k + x_k
        tt = root * (1 - root)
        den = s_k + ((d_k + d_k1 - 2 * s_k) * tt)
        num = s_k.pow(2) * (d_k1 * root.pow(2) + 2 * s_k * tt + d_k * (1 - root).pow(2))
        return out, -(torch.log(num) - 2 * torch.log(den))
    # spline_flow.py:163-179
    th = (v - x_k) / w_k
    tt = th * (1 - th)
    numer = h_k * (s_k * th.pow(2) + d_k * tt)
    den = s_k + ((d_k + d_k1 - 2 * s_k) * tt)
    out = y_k + numer / den
    num = s_k.pow(2) * (d_k1 * th.pow(2) + 2 * s_k * tt + d_k * (1 - th).pow(2))
    return out, torch.log(num) - 2 * torch.log(den)


def unconstrained_rqs(v, uw, uh, ud, inverse, bound):
    """spline_flow.py:29-68: identity tails outside [-bound, bound], RQS inside.

    Departure from the reference, on purpose: when *no* element lies inside the
    interval the reference crashes in RQS (torch.min of an empty tensor,
    spline_flow.py:85); the oracle returns the identity map.
    """
    inside = (v >= -bound) & (v <= bound)
    out = torch.zeros_like(v)
    lad = torch.zeros_like(v)
    ud = F.pad(ud, pad=(1, 1))
    const = math.log(math.exp(1 - MIN_D) - 1)  # spline_flow.py:47
    ud[..., 0] = const
    ud[..., -1] = const
    out[~inside] = v[~inside]
    if inside.any():
        out[inside], lad[inside] = rqs(
            v[inside], uw[inside, :], uh[inside, :], ud[inside, :], inverse, -bound, bound
        )
    return out, lad


def _spline_params(raw, K, B):
    """spline_flow.py:253-256 (and :208-211): split, first softmax x 2B, first softplus."""
    W, H, D = torch.split(raw, K, dim=-1)
    W, H = torch.softmax(W, dim=-1), torch.softmax(H, dim=-1)
    return 2 * B * W, 2 * B * H, F.softplus(D)


def nsf_cl(p, spec, v, inverse: bool):
    """spline_flow.py:249-285."""
    d, K, B = spec["dim"], spec["K"], spec["B"]
    h = d // 2
    ld = torch.zeros(v.shape[0], dtype=v.dtype)
    lower, upper = v[:, :h], v[:, h:]
    f1, f2 = sub(p, "f1."), sub(p, "f2.")
    if not inverse:
        W, H, D = _spline_params(mlp(f1, lower).reshape(-1, h, 3 * K - 1), K, B)
        upper, l = unconstrained_rqs(upper, W, H, D, False, B)
        ld += torch.sum(l, dim=1)
        W, H, D = _spline_params(mlp(f2, upper).reshape(-1, h, 3 * K - 1), K, B)
        lower, l = unconstrained_rqs(lower, W, H, D, False, B)
        ld += torch.sum(l, dim=1)
    else:
        W, H, D = _spline_params(mlp(f2, upper).reshape(-1, h, 3 * K - 1), K, B)
        lower, l = unconstrained_rqs(lower, W, H, D, True, B)
        ld += torch.sum(l, dim=1)
        W, H, D = _spline_params(mlp(f1, lower).reshape(-1, h, 3 * K - 1), K, B)
        upper, l = unconstrained_rqs(upper, W, H, D, True, B)
        ld += torch.sum(l, dim=1)
    return torch.cat([lower, upper], dim=1), ld


def nsf_ar(p, spec, v, inverse: bool):
    """spline_flow.py:199-235.  Note the convention: forward() runs the spline inverse."""
    d, K, B = spec["dim"], spec["K"], spec["B"]
    n = v.shape[0]
    out = torch.zeros_like(v)
    ld = torch.zeros(n, dtype=v.dtype)
    for i in range(d):
        if i == 0:
            raw = p["init_param"].expand(n, 3 * K - 1)
        else:
            # forward conditions on its own outputs, inverse on its inputs (:207 / :226)
            ctx = out[:, :i] if not inverse else v[:, :i]
            raw = mlp(sub(p, f"layers.{i - 1}."), ctx)
        W, H, D = _spline_params(raw, K, B)
        out[:, i], l = unconstrained_rqs(v[:, i], W, H, D, not inverse, B)
        ld += l
    return out, ld


# ---------------------------------------------------------------------------
# dispatch + containers
# ---------------------------------------------------------------------------
def apply_flow(p, spec, v, inverse: bool, mask=None):
    t = spec["type"]
    if t in ("AffineConstantFlow", "ActNormFlow"):
        return affine_constant(p, spec, v, inverse)
    if t == "AffineHalfFlow":
        return affine_half(p, spec, v, inverse)
    if t == "Glow":
        return glow(p, spec, v, inverse)
    if t == "MAF":
        return maf_density(p, spec, v) if inverse else maf_sample(p, spec, v)
    if t == "IAF":  # maf.py:70-72 swaps the two directions
        return maf_sample(p, spec, v) if inverse else maf_density(p, spec, v)
    if t == "NSF_CL":
        return nsf_cl(p, spec, v, inverse)
    if t == "NSF_AR":
        return nsf_ar(p, spec, v, inverse)
    if t == "RNVP":
        if inverse:
            raise NotImplementedError("RNVP has no inverse (rnvp.py)")
        return rnvp(p, spec, v, mask)
    raise ValueError(f"unknown flow type {t}")


def stack(sd: dict, specs: list[dict], v, inverse: bool, tape=None, prefix="flows."):
    """core.py:17-35 -- returns (list of intermediates incl. the input, log_det[B])."""
    ld = torch.zeros(v.size(0), dtype=v.dtype)
    outs = [v]
    order = list(enumerate(specs))
    if inverse:
        order = order[::-1]
    for i, spec in order:
        p = sub(sd, f"{prefix}{i}.")
        mask = tape.bernoulli(v.shape) if spec["type"] == "RNVP" else None
        v, l = apply_flow(p, spec, v, inverse, mask)
        ld = ld + l
        outs.append(v)
    return outs, ld


def std_normal_log_prob(z):
    """MultivariateNormal(0, I).log_prob as used by tests/test_flows.py:38 and core.py:49."""
    d = z.size(1)
    return -0.5 * (z * z).sum(1) - 0.5 * d * math.log(2 * math.pi)


def log_prob(sd, specs, x, prefix="flows."):
    """tests/test_flows.py:22-24: log_det of inverse() + base log-prob of the last z."""
    zs, ld = stack(sd, specs, x, True, prefix=prefix)
    return ld + std_normal_log_prob(zs[-1])
