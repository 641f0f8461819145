"""The reference's end-to-end training regressions (tests/test_flows.py:14-118) run on the drop-in modules:
same stacks, same data, same optimiser, same step counts and the SAME loss bounds.  Forward passes go through
mnf_flow_stack_run, gradients through mnf_flow_stack_backward."""

import pytest
import torch
from torch.distributions import MultivariateNormal

pytestmark = pytest.mark.gpu


def _train(model, optim, samples, steps=70):
    for _ in range(steps):
        _, log_det = model.inverse(samples)
        base_log_prob = model.base_log_prob(samples)
        loss = -(log_det + base_log_prob).sum()
        model.zero_grad()
        loss.backward()
        optim.step()
    return float(loss.detach())


def _stacks():
    import torch_mnf.flows as nf

    def with_actnorm(flows):
        out = []
        for f in flows:
            out += [nf.ActNormFlow(dim=2), f]
        return out

    return {
        "rnvp": (lambda: [nf.AffineHalfFlow(dim=2, parity=i % 2 == 0) for i in range(2)], 236),
        "maf": (lambda: [nf.MAF(dim=2, parity=i % 2 == 0) for i in range(2)], 250),
        "maf_actnorm": (lambda: with_actnorm([nf.MAF(dim=2, parity=i % 2 == 0) for i in range(2)]), 226),
        "iaf": (lambda: [nf.IAF(dim=2, parity=i % 2 == 0) for i in range(2)], 300),
        "glow": (lambda: [nf.Glow(dim=2) for _ in range(2)], 308),
        "glow_actnorm": (lambda: with_actnorm([nf.Glow(dim=2) for _ in range(2)]), 246),
        "nsfcl": (lambda: [nf.NSF_CL(dim=2, K=8, B=3, n_h=16) for _ in range(2)], 207),
        "nsfcl_actnorm": (lambda: with_actnorm([nf.NSF_CL(dim=2, K=8, B=3, n_h=16) for _ in range(2)]), 184),
        "nsfar": (lambda: [nf.NSF_AR(dim=2, K=8, B=3, n_h=16) for _ in range(2)], 318),
        "nsfar_actnorm": (lambda: with_actnorm([nf.NSF_AR(dim=2, K=8, B=3, n_h=16) for _ in range(2)]), 213),
    }


@pytest.mark.parametrize("name", ["rnvp", "maf", "maf_actnorm", "iaf", "glow", "glow_actnorm", "nsfcl",
                                  "nsfcl_actnorm", "nsfar", "nsfar_actnorm"])
def test_training_regression(name):
    import torch_mnf.flows as nf
    from torch_mnf import data

    make, bound = _stacks()[name]
    torch.manual_seed(0)
    samples = data.sample_moons(128).cuda()
    base = MultivariateNormal(torch.zeros(2), torch.eye(2))
    model = nf.NormalizingFlowModel(base, make()).cuda()
    adam = torch.optim.Adam(model.parameters())
    loss1 = _train(model, adam, samples, steps=1)
    loss2 = _train(model, adam, samples)
    assert loss1 > loss2
    assert loss2 < bound, f"{name}: loss {loss2:.4f} not below the reference's bound {bound}"
