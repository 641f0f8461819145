"""GPU parity of the tensor-core dim-2 spline kernel (csrc/flow_tc.cu, kernel variant 4): the conditioner MLPs run as
3xTF32 products on tcgen05, so the bar is the SAME fp32 criterion as the FFMA kernels (tests/test_flows_gpu.py), not the
2e-3 class of the single-pass TF32 GEMMs."""

import pytest
import torch

from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden, random_flow_sd, t
from tests.test_flows_gpu import ORACLE_CASES, close, close_vs_oracle

pytestmark = pytest.mark.gpu


def test_golden_nsfcl3_stack_on_tensor_cores():
    g = load_golden("nsfcl3_stack")
    sd, specs = golden_sd(g), golden_spec(g)
    prog = load_flow_model(specs, sd)._program()
    x = t(g, "inv/x").cuda()
    y, ld, inter, lp = prog.run(x, inverse=True, want_inter=True, want_base_lp=True, kernel=4)
    close(y, t(g, "inv/z"), "z", atol_scale=2e-5)
    close(ld, t(g, "inv/ld"), "log_det", atol_scale=4e-5)
    close(lp, t(g, "inv/base_log_prob"), "base_log_prob", atol_scale=4e-5)
    close(inter[-1], t(g, "inv/z"), "last intermediate", atol_scale=2e-5)
    close_vs_oracle((y, ld), sd, specs, t(g, "inv/x"), True, "nsfcl3_stack tc inverse")
    z = t(g, "fwd/z").cuda()
    y, ld, _, _ = prog.run(z, inverse=False, kernel=4)
    close(y, t(g, "fwd/x"), "x", atol_scale=2e-5)
    close(ld, t(g, "fwd/ld"), "log_det fwd", atol_scale=4e-5)
    # ragged: a batch that is not a multiple of the 128-point tile, and a single point
    for n in (101, 1):
        y, ld, _, _ = prog.run(x[:n].contiguous(), inverse=True, kernel=4)
        close(y, t(g, "inv/z")[:n], "z ragged", atol_scale=2e-5)
        close(ld, t(g, "inv/ld")[:n], "log_det ragged", atol_scale=4e-5)


@pytest.mark.parametrize("n_rows", [20001, 148 * 6 * 128 + 77])
def test_cfg2_stack_vs_oracle_on_tensor_cores(n_rows):
    """Seeded batch incl. tail points and exact +-B against the CPU oracle, both directions, every output."""
    specs = ORACLE_CASES["cfg2_shape"]
    sd = random_flow_sd(specs, seed=3, scale=0.6)
    prog = load_flow_model(specs, sd)._program()
    x = 1.5 * torch.randn(n_rows, 2, generator=torch.Generator().manual_seed(11))
    x[0, :], x[1, :], x[2, :] = 3.0, -3.0, 0.0
    for inverse in (True, False):
        y, ld, inter, lp = prog.run(x.cuda(), inverse=inverse, want_inter=True, want_base_lp=True, kernel=4)
        close_vs_oracle((y, ld), sd, specs, x, inverse, f"cfg2 tc inverse={inverse}")
        yg, ldg, interg, lpg = prog.run(x.cuda(), inverse=inverse, want_inter=True, want_base_lp=True, kernel="generic")
        torch.testing.assert_close(inter, interg, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(lp, lpg, rtol=1e-4, atol=3e-4)
    # fused log-prob mode (the bench call)
    _, _, _, lp1 = prog.run(x.cuda(), True, log_prob_only=True, kernel=4)
    torch.testing.assert_close(lp1, ldg_lp(prog, x.cuda()), rtol=1e-4, atol=5e-4)


def ldg_lp(prog, x):
    _, ld, _, lp = prog.run(x, True, want_base_lp=True, kernel="generic")
    return ld + lp


def test_other_shapes_fall_back():
    """Stacks outside the kernel's shape class (here K=5, n_h=8) still run when variant 4 is requested."""
    specs = ORACLE_CASES["nsf_default"]
    sd = random_flow_sd(specs, seed=1, scale=0.5)
    prog = load_flow_model(specs, sd)._program()
    x = torch.randn(1000, 2, generator=torch.Generator().manual_seed(1))
    y, ld, _, _ = prog.run(x.cuda(), inverse=True, kernel=4)
    close_vs_oracle((y, ld), sd, specs, x, True, "nsf_default via variant 4")


def test_activations_outside_fp16_range_take_the_exact_path():
    """The fp16-split operands overflow for |activation| > 65504: the kernel must notice (non-finite conditioner
    outputs) and re-evaluate those points' conditioners in fp32.  Conditioning coordinates of 3e5 .. 1e7 with the
    transformed coordinate inside [-B, B] exercise exactly that."""
    specs = [{"type": "NSF_CL", "dim": 2, "K": 8, "B": 3, "n_h": 16}] * 2
    sd = random_flow_sd(specs, seed=5, scale=0.6)
    prog = load_flow_model(specs, sd)._program()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4096, 2, generator=g)
    x[::7, 0] = 3e5 * (1 + 30 * torch.rand(x[::7].size(0), generator=g))  # f1 conditions on x[:, 0]
    x[3::11, 1] = -2e6
    for inverse in (True, False):
        y, ld, _, _ = prog.run(x.cuda(), inverse=inverse, kernel=4)
        yg, ldg, _, _ = prog.run(x.cuda(), inverse=inverse, kernel="generic")
        assert torch.isfinite(y).all() and torch.isfinite(ld).all()
        torch.testing.assert_close(y, yg, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(ld, ldg, rtol=1e-4, atol=3e-4)
