"""Pin the CPU oracle (oracle/) to outputs recorded from the real reference (tests/golden).

The oracle restates the reference with the same ATen ops in the same order, so on the same
torch build it must reproduce the reference to ~1 ulp; the tolerance below (1e-6 relative
with a 1e-6 absolute floor) allows for different intra-op threading only."""

import pytest
import torch

from oracle import flows_cpu, mnf_cpu
from tests.helpers import golden_sd, golden_spec, golden_tape, load_golden, t

FLOW_CASES = [
    "rnvp9_moons", "nsfcl3_stack", "nsfcl_d4", "nsfar2_d3", "maf9_d64", "maf3_d8",
    "maf_iaf_d2", "affine_misc_d4",
]


def close(a, b, rtol=1e-6, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", FLOW_CASES)
def test_flow_stack_matches_reference(name):
    g = load_golden(name)
    sd, specs = golden_sd(g), golden_spec(g)
    zs, ld = flows_cpu.stack(sd, specs, t(g, "inv/x"), inverse=True)
    assert len(zs) == len(specs) + 1
    close(zs[-1], t(g, "inv/z"))
    close(ld, t(g, "inv/ld"))
    close(zs[len(zs) // 2], t(g, "inv/z_mid"))
    close(flows_cpu.std_normal_log_prob(zs[-1]), t(g, "inv/base_log_prob"), rtol=1e-5, atol=1e-5)
    if "fwd/z" in g:
        xs, ld = flows_cpu.stack(sd, specs, t(g, "fwd/z"), inverse=False)
        close(xs[-1], t(g, "fwd/x"))
        close(ld, t(g, "fwd/ld"))
        close(xs[len(xs) // 2], t(g, "fwd/x_mid"))


def test_actnorm_init():
    g = load_golden("actnorm_init")
    spec = {"type": "ActNormFlow", "dim": 3, "scale": True, "shift": True}
    # any non-zero s/t triggers the init (affine_constant_flow.py:45,47)
    p = {"s": torch.ones(1, 3), "t": torch.ones(1, 3)}
    s, tt = flows_cpu.actnorm_init(p, spec, t(g, "x"))
    close(s, t(g, "s"))
    close(tt, t(g, "t"))
    z, ld = flows_cpu.affine_constant({"s": s, "t": tt}, spec, t(g, "x"), inverse=True)
    close(z, t(g, "z"))
    close(ld, t(g, "ld"))


def test_rnvp_mnf():
    g = load_golden("rnvp_mnf_d10")
    sd, specs = golden_sd(g), golden_spec(g)
    xs, ld = flows_cpu.stack(sd, specs, t(g, "z"), inverse=False, tape=golden_tape(g, "noise/"))
    close(xs[-1], t(g, "x"))
    close(ld, t(g, "ld"))


def test_made_masks_match_reference_buffers():
    g = load_golden("maf9_d64")
    sd = golden_sd(g)
    masks = flows_cpu.made_masks(64, [24, 24, 24], 128, natural=True)
    for i, m in enumerate(masks):
        ref = sd[f"flows.3.net.{2 * i}.mask"]
        assert ref.dtype == torch.bool
        assert torch.equal(torch.from_numpy(m), ref)


@pytest.mark.parametrize("name", ["mnf_linear_20x7", "mnf_linear_256x128"])
def test_mnf_linear(name):
    g = load_golden(name)
    sd = golden_sd(g)
    tape = golden_tape(g, "fwd_noise/")
    y = mnf_cpu.linear_forward(sd, t(g, "x"), tape)
    assert tape.pos == len(tape.draws)
    close(y, t(g, "fwd/y"), rtol=1e-5, atol=1e-5)
    tape = golden_tape(g, "kl_noise/")
    kl = mnf_cpu.linear_kl_div(sd, tape)
    assert tape.pos == len(tape.draws)
    close(kl, t(g, "kl/value"), rtol=1e-6, atol=1e-4)


def test_mnf_conv():
    g = load_golden("mnf_conv_2x3k3")
    sd = golden_sd(g)
    tape = golden_tape(g, "fwd_noise/")
    y = mnf_cpu.conv_forward(sd, t(g, "x"), tape)
    assert tape.pos == len(tape.draws)
    close(y, t(g, "fwd/y"), rtol=1e-5, atol=1e-5)
    tape = golden_tape(g, "kl_noise/")
    kl = mnf_cpu.conv_kl_div(sd, tape)
    assert tape.pos == len(tape.draws)
    close(kl, t(g, "kl/value"), rtol=1e-6, atol=1e-4)


def test_mnf_lenet():
    g = load_golden("mnf_lenet")
    sd = golden_sd(g)
    tape = golden_tape(g, "fwd_noise/")
    assert len(tape.draws) == 16  # 16 draws per MNF-LeNet forward
    y = mnf_cpu.lenet_forward(sd, t(g, "x"), tape)
    close(y, t(g, "fwd/y"), rtol=1e-5, atol=1e-5)
    tape = golden_tape(g, "kl_noise/")
    kl = mnf_cpu.lenet_kl_div(sd, tape)
    assert tape.pos == len(tape.draws)
    close(kl, t(g, "kl/value"), rtol=1e-6, atol=1e-2)


def test_spline_all_outside_is_identity():
    """Reference crashes here (spline_flow.py:85 on an empty tensor); oracle maps to identity."""
    v = torch.tensor([[4.0], [-5.0]])
    W = torch.zeros(2, 1, 8)
    out, lad = flows_cpu.unconstrained_rqs(v, W, W.clone(), torch.zeros(2, 1, 7), False, 3)
    assert torch.equal(out, v) and torch.equal(lad, torch.zeros_like(v))
