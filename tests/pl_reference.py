"""CPU restatement (numpy, fp64) of the table builder of csrc/flow_pl.cu -- TEST INFRASTRUCTURE, like oracle/: imported by
tests only.  A Linear / LeakyReLU(0.2) MLP (models/mlp.py:4-12) of ONE scalar input is piecewise linear; `breakpoints`
finds the kinks layer by layer (on every piece of the current partition each pre-activation of the next layer is affine,
so it has at most one zero there), `tables` returns per piece the affine map of every output."""

import numpy as np

SLOPE = float(np.float32(0.2))  # LeakyReLU(0.2) on fp32 tensors


def piece_points(bp):
    """One point strictly inside each of the len(bp) + 1 pieces (same rule as flow_pl.cu::piece_point)."""
    if len(bp) == 0:
        return np.zeros(1), np.array([-np.inf]), np.array([np.inf])
    lo = np.concatenate([[-np.inf], bp])
    hi = np.concatenate([bp, [np.inf]])
    m = 0.5 * lo + 0.5 * hi
    m[0] = hi[0] - 1.0 - abs(hi[0])
    m[-1] = lo[-1] + 1.0 + abs(lo[-1])
    return m, lo, hi


def pre_maps(Ws, bs, upto, m):
    """(alpha, beta): pre-activations of Linear layer `upto` (0-based) are alpha * c + beta on the piece containing m."""
    a, b = np.ones(1), np.zeros(1)
    for t in range(upto + 1):
        al, be = Ws[t] @ a, Ws[t] @ b + bs[t]
        if t == upto:
            return al, be
        s = np.where(al * m + be > 0, 1.0, SLOPE)
        a, b = s * al, s * be


def breakpoints(nets):
    """Sorted kinks of a group of nets that read the same scalar (an AffineHalfFlow's s- and t-net share one table)."""
    bp = np.zeros(0)
    n_hidden = len(nets[0][0]) - 1
    for l in range(1, n_hidden + 1):
        m, lo, hi = piece_points(bp)
        new = []
        for i in range(len(m)):
            for Ws, bs in nets:
                al, be = pre_maps(Ws, bs, l - 1, m[i])
                with np.errstate(divide="ignore", invalid="ignore"):
                    tz = -be / al
                new += list(tz[(al != 0) & (tz > lo[i]) & (tz < hi[i]) & (np.abs(tz) < 3.0e38)])  # kinks inside the fp32 range
        bp = np.sort(np.concatenate([bp, new]))
    return bp


def tables(nets, bp):
    """Per piece and net: (A, B) with out = A * c + B."""
    m, _, _ = piece_points(bp)
    return [[pre_maps(Ws, bs, len(Ws) - 1, mi) for Ws, bs in nets] for mi in m]


def mlp(Ws, bs, c):
    """Layer-by-layer evaluation at the scalars c (fp64)."""
    h = np.asarray(c, dtype=np.float64)[None, :]
    for t, (W, b) in enumerate(zip(Ws, bs)):
        h = W @ h + b[:, None]
        if t + 1 < len(Ws):
            h = np.where(h > 0, h, SLOPE * h)
    return h


def nets_of_flow(sd, i, spec):
    """[(Ws, bs)] groups of flow i of a state_dict: NSF_CL -> [f1], [f2]; AffineHalfFlow -> [s_net, t_net] together."""
    names = ("f1", "f2") if spec["type"] == "NSF_CL" else ("s_net", "t_net")
    out = []
    for nm in names:
        ks = sorted((k for k in sd if k.startswith(f"flows.{i}.{nm}.") and k.endswith("weight")), key=lambda k: int(k.split(".")[-2]))
        if ks:
            out.append(([sd[k].double().numpy() for k in ks], [sd[k.replace("weight", "bias")].double().numpy() for k in ks]))
    return [[n] for n in out] if spec["type"] == "NSF_CL" else [out]
