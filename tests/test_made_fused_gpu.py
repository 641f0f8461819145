"""The fused tcgen05 MADE-chain kernel (csrc/made_fused.cu, mnf_made_density_fused): MAF.inverse for a stack of dim-64
MAF flows in one persistent kernel.  Parity against the reference's recorded outputs (tests/golden/maf9_d64.npz) and
against the CPU oracle on seeded stacks, at the rtol 2e-3 BASELINE.json states for the MADE tensor-core GEMMs (absolute
floor 2e-3 x RMS of the reference values); structural cases: partial last tile, fewer tiles than SMs, parity mixes
(the flips are folded into the packed weights), 1..4 hidden layers, intermediates, fused log-prob, every kernel variant
(2 / 3 tiles in flight x 1 / 2 threads per row)."""

import pytest
import torch

from oracle import flows_cpu
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden, random_flow_sd, t

pytestmark = pytest.mark.gpu


def tc_close(got, ref, what, extra_atol=0.0):
    ref = ref.float()
    torch.testing.assert_close(got.detach().float().cpu(), ref, rtol=2e-3, atol=2e-3 * float(ref.pow(2).mean().sqrt()) + extra_atol,
                               msg=lambda m: f"{what}: {m}")


def maf_specs(parities, h_sizes):
    return [{"type": "MAF", "dim": 64, "parity": bool(p), "h_sizes": list(h_sizes)} for p in parities]


def test_fused_kernel_vs_reference_golden():
    from torch_mnf.flows import maf

    g = load_golden("maf9_d64")
    model = load_flow_model(golden_spec(g), golden_sd(g))
    for f in model.flows:
        f.precision = "tf32"
    x = t(g, "inv/x").cuda()
    assert maf._tc_mode(list(model.flows)[::-1], x) == "fused"
    zs, ld = model.inverse(x)  # 96 rows: one partial tile, intermediates through per-flow launches
    assert len(zs) == 10
    tc_close(zs[-1], t(g, "inv/z"), "z")
    tc_close(ld, t(g, "inv/ld"), "log_det", extra_atol=2e-3)
    tc_close(zs[5], t(g, "inv/z_mid"), "z_mid")
    model.return_intermediates = False
    zs2, ld2 = model.inverse(x)  # one launch for the whole stack, flips folded into the weights
    assert len(zs2) == 2
    tc_close(zs2[-1], t(g, "inv/z"), "z chained")
    tc_close(ld2, t(g, "inv/ld"), "log_det chained", extra_atol=2e-3)
    lp = model.log_prob(x)
    tc_close(lp, t(g, "inv/ld") + t(g, "inv/base_log_prob"), "log_prob", extra_atol=2e-3)
    z1, ld1 = model.flows[8].inverse(x)  # single-flow module API
    tc_close(z1, zs[1].cpu(), "single flow")


@pytest.mark.parametrize("parities,h_sizes,n_rows", [
    ((0,), (24, 24, 24), 1000), ((1,), (24, 24, 24), 128), ((1, 1), (24, 24, 24), 300), ((0, 1, 0, 1, 0, 1, 0, 1, 0), (24, 24, 24), 5000),
    ((1, 0, 1), (31,), 777), ((0, 1), (16, 20), 4097), ((1, 0, 0, 1), (8, 31, 5, 17), 2500),
    (tuple(i % 2 for i in range(16)), (24, 24, 24), 1500),
])
@pytest.mark.parametrize("tiles", [21, 31, 41, 32])
def test_fused_kernel_vs_oracle(parities, h_sizes, n_rows, tiles):
    from torch_mnf.layers import made

    specs = maf_specs(parities, h_sizes)
    sd = random_flow_sd(specs, seed=len(parities) * 7 + n_rows, scale=0.7)
    for k in sd:  # keep exp(s) tame through deep stacks
        if k.endswith("net.%d.weight" % (2 * len(h_sizes))) or k.endswith("net.%d.bias" % (2 * len(h_sizes))):
            sd[k] = sd[k] * 0.3
    model = load_flow_model(specs, sd)
    for f in model.flows:
        f.precision = "tf32"
    x = torch.randn(n_rows, 64, generator=torch.Generator().manual_seed(n_rows))
    ref, ref_ld = flows_cpu.stack(sd, specs, x, inverse=True)
    made.VARIANT = tiles
    try:
        zs, ld = model.inverse(x.cuda())
        assert len(zs) == len(specs) + 1
        for i in range(1, len(zs)):
            tc_close(zs[i], ref[i], f"intermediate {i}")
        tc_close(ld, ref_ld, "log_det", extra_atol=2e-3)
        model.return_intermediates = False
        zs2, ld2 = model.inverse(x.cuda())
        tc_close(zs2[-1], ref[-1], "z chained")
        tc_close(ld2, ref_ld, "log_det chained", extra_atol=2e-3)
        lp = model.log_prob(x.cuda())
        tc_close(lp, flows_cpu.log_prob(sd, specs, x), "log_prob", extra_atol=2e-3)
        again, _ = model.inverse(x.cuda())
        assert torch.equal(again[-1], zs2[-1])  # deterministic
    finally:
        made.VARIANT = 0


def test_fused_kernel_full_batch_properties():
    """BASELINE config-3 size (2^20 rows): every row of the result against the exact-fp32 kernel of the same library at
    the tensor-core tolerance, two 4096-row windows against the oracle, log-det consistency of the fused log-prob."""
    g = load_golden("maf9_d64")
    sd, specs = golden_sd(g), golden_spec(g)
    model = load_flow_model(specs, sd, return_intermediates=False)
    n = 1 << 20
    x = torch.randn(n, 64, generator=torch.Generator().manual_seed(0))
    xd = x.cuda()
    zs, ld = model.inverse(xd)  # "auto": fused kernel at this size
    for win in (slice(0, 4096), slice(n - 4096, n)):
        ref, ref_ld = flows_cpu.stack(sd, specs, x[win], inverse=True)
        tc_close(zs[-1][win], ref[-1], f"z rows {win}")
        tc_close(ld[win], ref_ld, f"log_det rows {win}", extra_atol=2e-3)
    for f in model.flows:
        f.precision = "fp32"
    zs32, ld32 = model.inverse(xd)
    scale = float(zs32[-1].pow(2).mean().sqrt())
    err = (zs[-1] - zs32[-1]).abs()
    assert float((err > 2e-3 * zs32[-1].abs() + 2e-3 * scale).float().mean()) == 0.0
    assert float((ld - ld32).abs().max()) < 2e-3 * float(ld32.pow(2).mean().sqrt()) + 2e-3
    for f in model.flows:
        f.precision = "auto"
    lp = model.log_prob(xd)
    ref_lp = ld32 - 0.5 * zs32[-1].square().sum(1) - 32 * 1.8378770664093453
    assert float((lp - ref_lp).abs().max()) < 2e-3 * float(ref_lp.pow(2).mean().sqrt()) + 2e-3
    assert torch.isfinite(lp).all()


def test_small_batches_keep_the_exact_path():
    from torch_mnf.flows import maf

    g = load_golden("maf9_d64")
    model = load_flow_model(golden_spec(g), golden_sd(g))
    x = t(g, "inv/x").cuda()
    assert maf._tc_mode(list(model.flows), x) is None  # 96 rows under "auto": exact fp32 kernel
    zs, ld = model.inverse(x)
    torch.testing.assert_close(zs[-1].cpu(), t(g, "inv/z"), rtol=1e-5, atol=2e-5)
