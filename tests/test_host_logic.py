"""Host-side logic that needs no GPU: drop-in API surface, state_dict compatibility with the
reference (golden fixtures hold reference state_dicts), flow-program packing, kernel planning."""

import pytest
import torch

from oracle import flows_cpu
from tests.helpers import build_flow, golden_sd, golden_spec, load_flow_model, load_golden

CASES = ["rnvp9_moons", "nsfcl3_stack", "nsfcl_d4", "nsfar2_d3", "maf9_d64", "maf3_d8", "maf_iaf_d2",
         "affine_misc_d4"]


@pytest.mark.parametrize("name", CASES)
def test_reference_state_dict_loads_strict(name):
    g = load_golden(name)
    model = load_flow_model(golden_spec(g), golden_sd(g), device="cpu")
    keys = {k for k in golden_sd(g) if not k.endswith(".P")}
    assert set(model.state_dict().keys()) == keys


def test_public_names_match_reference():
    import torch_mnf.flows as nf
    import torch_mnf.layers as L
    import torch_mnf.models as M

    for n in ["ActNormFlow", "AffineConstantFlow", "AffineHalfFlow", "NormalizingFlow", "NormalizingFlowModel",
              "Glow", "IAF", "MAF", "RNVP", "NSF_AR", "NSF_CL"]:  # flows/__init__.py:17-23
        assert hasattr(nf, n)
    for n in ["MADE", "MaskedLinear", "MNFConv2d", "MNFLinear"]:  # layers/__init__.py:1-3
        assert hasattr(L, n)
    for n in ["MLP", "MNFFeedForward", "MNFLeNet"]:
        assert hasattr(M, n)


def test_made_masks_equal_oracle_and_reference():
    from torch_mnf.layers import MADE

    made = MADE(64, [24, 24, 24], 128, natural_ordering=True)
    ref = golden_sd(load_golden("maf9_d64"))
    for i, m in enumerate(flows_cpu.made_masks(64, [24, 24, 24], 128, natural=True)):
        assert torch.equal(made[2 * i].mask, torch.from_numpy(m))
        assert torch.equal(made[2 * i].mask, ref[f"flows.0.net.{2 * i}.mask"])
    made = MADE(6, [7, 5], 12, num_masks=3)  # random ordering + mask cycling keeps the autoregressive property
    for _ in range(4):
        made.update_masks()
        m = [layer.mask.float() for layer in made if hasattr(layer, "mask")]
        conn = (m[0] @ m[1] @ m[2])[:, :6]  # input i -> output j path count
        order = torch.from_numpy(made.m[-1])
        for i in range(6):
            for j in range(6):
                if order[i] >= order[j]:
                    assert conn[i, j] == 0


def test_flow_program_packing_and_plan():
    from torch_mnf import _lib
    from torch_mnf._program import FlowProgram

    specs = [{"type": "AffineHalfFlow", "dim": 2, "parity": bool(i % 2), "scale": True, "shift": True,
              "h_sizes": [24, 24, 24]} for i in range(9)]
    flows = [build_flow(s) for s in specs]
    prog = FlowProgram(flows)
    prog._build(torch.device("cpu"))
    assert prog._n_ops == 9 and prog._blob.numel() % 4 == 0
    per_net = 24 + 24 + 2 * (24 * 24 + 24) + 24 + 1
    assert prog._blob.numel() >= 18 * per_net
    for k in range(9):
        op = prog._ops[k]
        assert op.type == _lib.OP_AFFINE_HALF and op.n_lin == 4 and list(op.sizes[:5]) == [1, 24, 24, 24, 1]
        assert op.net_off[0] % 4 == 0 and op.net_off[1] % 4 == 0
        assert bool(op.flags & _lib.FLAG_PARITY) == bool(k % 2)
        w0 = flows[k].s_net[0].weight.detach().reshape(-1)
        assert torch.equal(prog._blob[op.net_off[0]: op.net_off[0] + 24], w0)
    assert prog.plan(torch.device("cpu"), 2) == 1  # BASELINE config 1 -> register-resident kernel
    key = prog._key
    prog._build(torch.device("cpu"))
    assert prog._key is key  # cached
    with torch.no_grad():
        flows[3].t_net[2].bias.add_(1.0)
    prog._build(torch.device("cpu"))
    assert prog._key != key  # parameter change invalidates the packed blob
    wide = FlowProgram([build_flow({"type": "NSF_CL", "dim": 4, "K": 5, "B": 3, "n_h": 8})])
    assert wide.plan(torch.device("cpu"), 4) == 0  # dim 4 -> generic interpreter


def test_maf_masks_are_folded_into_packed_weights():
    from torch_mnf._program import FlowProgram

    f = build_flow({"type": "MAF", "dim": 8, "parity": True, "h_sizes": [16, 16]})
    prog = FlowProgram([f])
    prog._build(torch.device("cpu"))
    lin0 = f.net[0]
    packed = prog._blob[prog._ops[0].net_off[0]: prog._ops[0].net_off[0] + 16 * 8].view(16, 8)
    assert torch.equal(packed, (lin0.weight * lin0.mask.float().T).detach())


def test_shard_range_covers_rows_exactly():
    from torch_mnf.distributed import shard_range

    for n in (0, 1, 7, 500, 1 << 24):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
