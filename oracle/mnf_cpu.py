"""CPU oracle: MNFLinear / MNFConv2d forward + kl_div and the MNF-LeNet composition.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Noise is taken from a tape
(oracle/noise.py) in the reference's draw order:

* MNFLinear.forward : normal[R,n_in] -> bernoulli[R,n_in] per q-RNVP -> normal[R,n_out]
* MNFLinear.kl_div  : normal[1,n_in] -> bernoulli[1,n_in] x n_q -> normal[n_out,n_in]
                      -> bernoulli[1,n_in] x n_r
* MNFConv2d.forward : normal[n_out] -> bernoulli[1,n_out] x n_q -> normal[R,n_out,H',W']
* MNFConv2d.kl_div  : normal[n_out] -> bernoulli[1,n_out] x n_q -> normal[n_in*k*k]
                      -> normal[] -> bernoulli[1,n_out] x n_r
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from .flows_cpu import rnvp, sub


def _n_flows(p: dict, name: str) -> int:
    ids = {int(k.split(".")[2]) for k in p if k.startswith(f"{name}.flows.")}
    return len(ids)


def _rnvp_stack(p, name, z, tape):
    """core.py:17-25 over RNVP flows; returns (z_T, log_det[B])."""
    ld = torch.zeros(z.size(0))
    for i in range(_n_flows(p, name)):
        mask = tape.bernoulli(z.shape)
        z, l = rnvp(sub(p, f"{name}.flows.{i}."), {"type": "RNVP"}, z, mask)
        ld = ld + l
    return z, ld


# ---------------------------------------------------------------------------
# MNFLinear
# ---------------------------------------------------------------------------
def linear_sample_z(p, batch, tape):
    """mnf_linear.py:58-64."""
    std = p["q0_log_var"].exp().sqrt().repeat(batch, 1)
    eps = tape.normal(std.shape)
    z = p["q0_mean"] + std * eps
    zT, ld = _rnvp_stack(p, "flow_q", z, tape)
    return zT, ld.squeeze()


def linear_forward(p, x, tape):
    """mnf_linear.py:46-56."""
    z, _ = linear_sample_z(p, x.size(0), tape)
    mean = x * z @ p["W_mean"].T + p["b_mean"]
    var = x**2 @ p["W_log_var"].exp().T + p["b_log_var"].exp()
    eps = tape.normal(var.shape)
    return mean + var.sqrt() * eps


def linear_kl_div(p, tape):
    """mnf_linear.py:66-90."""
    z, ld_q = linear_sample_z(p, 1, tape)
    W_mean = z * p["W_mean"]
    W_var = p["W_log_var"].exp()
    eps_w = tape.normal(W_var.shape)
    weight = W_mean + W_var.sqrt() * eps_w
    kl_W = 0.5 * torch.sum(-W_var.log() + W_var + W_mean**2 - 1)
    kl_b = 0.5 * torch.sum(-p["b_log_var"] + p["b_log_var"].exp() + p["b_mean"] ** 2 - 1)
    log_q = -ld_q - 0.5 * p["q0_log_var"].sum()
    act = torch.tanh(p["r0_c"] @ weight.T)
    mean_r = p["r0_b1"].ger(act).mean(1)
    log_var_r = p["r0_b2"].ger(act).mean(1)
    zT, ld_r = _rnvp_stack(p, "flow_r", z, tape)
    (ld_r,) = ld_r
    log_r = ld_r + 0.5 * torch.sum(-log_var_r.exp() * (zT - mean_r) ** 2 + log_var_r)
    return kl_W + kl_b + log_q - log_r


# ---------------------------------------------------------------------------
# MNFConv2d
# ---------------------------------------------------------------------------
def conv_sample_z(p, tape):
    """mnf_conv.py:80-88."""
    std = p["q0_log_var"].exp().sqrt()
    eps = tape.normal(std.shape)
    z = p["q0_mean"] + std * eps
    zT, ld = _rnvp_stack(p, "flow_q", z[None, ...], tape)
    return zT, ld.squeeze()


def conv_forward(p, x, tape):
    """mnf_conv.py:67-78 (b_mean is the all-zero plain tensor of mnf_conv.py:45)."""
    z, _ = conv_sample_z(p, tape)
    n_out = p["W_mean"].shape[0]
    W_mean = p["W_mean"] * z.view(-1, 1, 1, 1)
    mean = F.conv2d(x, weight=W_mean, bias=torch.zeros(n_out, dtype=W_mean.dtype))
    var = F.conv2d(x**2, weight=p["W_log_var"].exp(), bias=p["b_log_var"].exp())
    eps = tape.normal(var.shape)
    return mean + var.sqrt() * eps


def conv_kl_div(p, tape):
    """mnf_conv.py:90-133."""
    z, ld_q = conv_sample_z(p, tape)
    n_out = p["W_mean"].shape[0]
    W_var = p["W_log_var"].exp()
    b_var = p["b_log_var"].exp()
    W_mean = p["W_mean"] * z.view(-1, 1, 1, 1)
    b_mean = torch.zeros(n_out) * z
    kl_W = 0.5 * torch.sum(-W_var.log() + W_var + W_mean**2 - 1)
    kl_b = 0.5 * torch.sum(-b_var.log() + b_var + b_mean**2 - 1)
    log_q = -ld_q - 0.5 * p["q0_log_var"].sum()
    Wm = W_mean.view(-1, n_out) @ p["r0_c"]  # memory reinterpretation, mnf_conv.py:107
    Ws = W_var.sqrt().view(-1, n_out) @ p["r0_c"]
    eps_w = tape.normal(Ws.shape)
    act = Wm + Ws * eps_w
    bm = torch.sum(b_mean * p["r0_c"])
    bv = torch.sum(p["b_log_var"].exp() * p["r0_c"] ** 2)
    eps_b = tape.normal(())
    act = act + (bm + bv.sqrt() * eps_b)
    mean_r = p["r0_b1"].ger(act).mean(1)
    log_var_r = p["r0_b2"].ger(act).mean(1)
    zT, ld_r = _rnvp_stack(p, "flow_r", z, tape)
    (ld_r,) = ld_r
    log_r = ld_r + 0.5 * torch.sum(-log_var_r.exp() * (zT - mean_r) ** 2 + log_var_r)
    return kl_W + kl_b + log_q - log_r


# ---------------------------------------------------------------------------
# MNF-LeNet (models/mnf_lenet.py:13-32)
# ---------------------------------------------------------------------------
LENET_MNF_SLOTS = ("0.", "3.", "7.", "9.")  # nn.Sequential indices of the MNF layers


def lenet_forward(sd, x, tape):
    h = conv_forward(sub(sd, "0."), x, tape)
    h = F.max_pool2d(torch.relu(h), 2)
    h = conv_forward(sub(sd, "3."), h, tape)
    h = F.max_pool2d(torch.relu(h), 2)
    h = h.flatten(1)
    h = torch.relu(linear_forward(sub(sd, "7."), h, tape))
    h = linear_forward(sub(sd, "9."), h, tape)
    return F.log_softmax(h, dim=-1)


def lenet_kl_div(sd, tape):
    return (
        conv_kl_div(sub(sd, "0."), tape)
        + conv_kl_div(sub(sd, "3."), tape)
        + linear_kl_div(sub(sd, "7."), tape)
        + linear_kl_div(sub(sd, "9."), tape)
    )
