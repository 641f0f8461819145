"""Generate tests/golden/*.npz from the REAL reference (build container only).

TEST INFRASTRUCTURE.  Run as

    python oracle/make_golden.py            # needs /root/reference (read-only checkout)

in a process that has the reference's ``torch_mnf`` on ``sys.path`` and *not* this
repo's drop-in package of the same name.  matplotlib / seaborn are not installed in
the image; the reference only touches them inside plotting helpers, so two empty stub
modules are injected.  Each fixture stores: the flow ``spec`` (JSON), the reference
``state_dict``, the inputs, the recorded noise tape, and the reference's outputs.
Fixtures are deliberately small (the whole directory is a few MB); the GPU parity tests
use them directly and use the oracle (pinned to them by tests/test_oracle_golden.py)
for larger seeded cases.
"""

from __future__ import annotations

import json
import os
import sys
import types

REF = os.environ.get("MNF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
sys.path.insert(1, os.path.dirname(HERE))  # for `oracle.noise` only

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch_mnf.flows as nf  # noqa: E402  (the reference)
from sklearn.datasets import make_moons  # noqa: E402
from torch.distributions import MultivariateNormal  # noqa: E402
from torch_mnf.layers import MNFConv2d, MNFLinear  # noqa: E402
from torch_mnf.models import MNFLeNet  # noqa: E402

from oracle.noise import NoiseTape, record  # noqa: E402

assert os.path.realpath(nf.__file__).startswith(os.path.realpath(REF)), nf.__file__


def moons(n):  # torch_mnf/data.py:21-24
    return torch.as_tensor(make_moons(n, noise=0.05, random_state=0)[0]).float()


def flow_spec(f) -> dict:
    t = type(f).__name__
    if t in ("AffineConstantFlow", "ActNormFlow"):
        return {
            "type": t,
            "dim": f.s.shape[1],
            "scale": isinstance(f.s, torch.nn.Parameter),
            "shift": isinstance(f.t, torch.nn.Parameter),
        }
    if t == "AffineHalfFlow":
        s_ok = isinstance(f.s_net, torch.nn.Module)
        t_ok = isinstance(f.t_net, torch.nn.Module)
        net = f.s_net if s_ok else f.t_net
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        return {
            "type": t,
            "dim": 2 * lin[0].in_features,
            "parity": bool(f.parity),
            "scale": s_ok,
            "shift": t_ok,
            "h_sizes": [m.out_features for m in lin[:-1]],
        }
    if t == "Glow":
        return {"type": t, "dim": len(f.S)}
    if t in ("MAF", "IAF"):
        return {
            "type": t,
            "dim": f.net.n_in,
            "parity": bool(f.parity),
            "h_sizes": list(f.net.hidden_sizes),
        }
    if t in ("NSF_CL", "NSF_AR"):
        net = f.f1 if t == "NSF_CL" else f.layers[0]
        return {"type": t, "dim": f.dim, "K": f.K, "B": f.B, "n_h": net[0].out_features}
    if t == "RNVP":
        return {"type": t, "dim": f.t.out_features, "h_sizes": [f.net[0].out_features]}
    raise ValueError(t)


def sd_np(module, extra=None) -> dict:
    out = {f"sd/{k}": v.detach().cpu().numpy() for k, v in module.state_dict().items()}
    for k, v in (extra or {}).items():
        out[f"sd/{k}"] = v.detach().cpu().numpy()
    return out


def glow_extras(model) -> dict:
    """Glow.P is a plain attribute, not in state_dict (glow.py:14)."""
    return {f"flows.{i}.P": f.P for i, f in enumerate(model.flows) if isinstance(f, nf.Glow)}


def train(model, x, steps):
    """tests/test_flows.py:14-31."""
    if steps == 0:
        with torch.no_grad():
            model.inverse(x)  # triggers ActNorm data-dependent init
        return
    opt = torch.optim.Adam(model.parameters())
    for _ in range(steps):
        _, ld = model.inverse(x)
        loss = -(ld + model.base_log_prob(x)).sum()
        model.zero_grad()
        loss.backward()
        opt.step()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def flow_case(name, flows, dim, x_train, steps, x_eval, z_eval, do_forward=True):
    base = MultivariateNormal(torch.zeros(dim), torch.eye(dim))
    model = nf.NormalizingFlowModel(base, flows)
    train(model, x_train, steps)
    arrays = {"spec": np.array(json.dumps([flow_spec(f) for f in model.flows]))}
    arrays.update(sd_np(model, glow_extras(model)))
    with torch.no_grad():
        zs, ld = model.inverse(x_eval)
        arrays["inv/x"] = x_eval.numpy()
        arrays["inv/z"] = zs[-1].numpy()
        arrays["inv/ld"] = ld.numpy()
        arrays["inv/z_mid"] = zs[len(zs) // 2].numpy()
        arrays["inv/base_log_prob"] = model.base_log_prob(x_eval).numpy()
        if do_forward:
            xs, ld = model.forward(z_eval)
            arrays["fwd/z"] = z_eval.numpy()
            arrays["fwd/x"] = xs[-1].numpy()
            arrays["fwd/ld"] = ld.numpy()
            arrays["fwd/x_mid"] = xs[len(xs) // 2].numpy()
    save(name, **arrays)


def gen_flows():
    g = torch.Generator().manual_seed(1234)

    torch.manual_seed(0)
    flows = [nf.AffineHalfFlow(dim=2, parity=bool(i % 2)) for i in range(9)]
    x = torch.cat([moons(192), moons(64) + 0.03 * torch.randn(64, 2, generator=g)])  # in-distribution: exp(s) stays O(1)
    flow_case("rnvp9_moons", flows, 2, moons(128), 70, x, torch.randn(256, 2, generator=g))

    torch.manual_seed(0)
    flows = sum(
        [[nf.ActNormFlow(2), nf.Glow(2), nf.NSF_CL(2, K=8, B=3, n_h=16)] for _ in range(3)], []
    )
    x = 1.5 * torch.randn(509, 2, generator=g)
    x[:4] = torch.tensor([[3.0, -3.0], [-3.0, 3.0], [0.0, 0.0], [7.5, -9.0]])
    flow_case("nsfcl3_stack", flows, 2, moons(128), 40, x, 1.5 * torch.randn(384, 2, generator=g))

    torch.manual_seed(0)
    flows = [nf.NSF_CL(4, K=5, B=2, n_h=8), nf.NSF_CL(4, K=5, B=2, n_h=8)]
    x = 1.2 * torch.randn(200, 4, generator=g)
    flow_case("nsfcl_d4", flows, 4, x, 10, x, 1.2 * torch.randn(130, 4, generator=g))

    torch.manual_seed(0)
    flows = sum([[nf.ActNormFlow(3), nf.Glow(3), nf.NSF_AR(3, K=5, B=3, n_h=8)] for _ in range(2)], [])
    x = 1.5 * torch.randn(257, 3, generator=g)
    flow_case("nsfar2_d3", flows, 3, x, 20, x, 1.5 * torch.randn(129, 3, generator=g))

    torch.manual_seed(0)
    flows = [nf.MAF(dim=64, parity=bool(i % 2)) for i in range(9)]
    x = torch.randn(96, 64, generator=g)
    flow_case("maf9_d64", flows, 64, x, 5, x, None, do_forward=False)

    torch.manual_seed(0)
    flows = [nf.MAF(dim=8, parity=bool(i % 2), h_sizes=(16, 16)) for i in range(3)]
    x = torch.randn(77, 8, generator=g)
    flow_case("maf3_d8", flows, 8, x, 10, x, torch.randn(40, 8, generator=g))

    torch.manual_seed(0)
    flows = [nf.ActNormFlow(2), nf.MAF(2, parity=True), nf.ActNormFlow(2), nf.IAF(2, parity=False)]
    flow_case("maf_iaf_d2", flows, 2, moons(128), 30, moons(100), torch.randn(90, 2, generator=g))

    torch.manual_seed(0)
    flows = [
        nf.AffineConstantFlow(4, scale=False),
        nf.AffineHalfFlow(4, parity=False, scale=False),
        nf.AffineConstantFlow(4, shift=False),
        nf.AffineHalfFlow(4, parity=True, h_sizes=(8,)),
        nf.Glow(4),
        nf.AffineHalfFlow(4, parity=False, shift=False, h_sizes=(12, 6)),
        nf.AffineConstantFlow(4),
    ]
    x = torch.randn(150, 4, generator=g)
    flow_case("affine_misc_d4", flows, 4, x, 15, x, torch.randn(64, 4, generator=g))

    # ActNorm data-dependent init on its own (affine_constant_flow.py:42-50)
    torch.manual_seed(0)
    f = nf.ActNormFlow(3)
    x = 2.0 * torch.randn(300, 3, generator=g) + torch.tensor([1.0, -2.0, 0.5])
    with torch.no_grad():
        z, ld = f.inverse(x)
    save(
        "actnorm_init",
        x=x.numpy(), z=z.numpy(), ld=ld.numpy(), s=f.s.detach().numpy(), t=f.t.detach().numpy(),
    )

    # RNVP (MNF-style, random mask) on its own
    torch.manual_seed(0)
    f = nf.RNVP(10, h_sizes=(50,))
    z = torch.randn(37, 10, generator=g)
    tape = NoiseTape()
    with torch.no_grad(), record(tape):
        xo, ld = f.forward(z)
    save(
        "rnvp_mnf_d10",
        spec=np.array(json.dumps([flow_spec(f)])),
        z=z.numpy(), x=xo.numpy(), ld=ld.numpy(),
        **{f"sd/flows.0.{k}": v.numpy() for k, v in f.state_dict().items()},
        **tape.to_npz_dict(),
    )


def perturb_mnf(layer, g):
    """Default init has sigma^2 = e^-9; widen so the variance path matters in the tests."""
    with torch.no_grad():
        layer.W_log_var += 5.0 + 0.5 * torch.randn(layer.W_log_var.shape, generator=g)
        layer.b_log_var += 6.0
        layer.q0_log_var += 7.0
        if isinstance(layer.b_mean, torch.nn.Parameter):
            layer.b_mean += 0.3 * torch.randn(layer.b_mean.shape, generator=g)


def mnf_layer_case(name, layer, x):
    arrays = sd_np(layer)
    arrays["x"] = x.numpy()
    with torch.no_grad():
        t1 = NoiseTape()
        with record(t1):
            y = layer(x)
        arrays["fwd/y"] = y.numpy()
        arrays.update(t1.to_npz_dict("fwd_noise/"))
        t2 = NoiseTape()
        with record(t2):
            kl = layer.kl_div()
        arrays["kl/value"] = kl.numpy()
        arrays.update(t2.to_npz_dict("kl_noise/"))
    save(name, **arrays)


def gen_mnf():
    g = torch.Generator().manual_seed(4321)
    torch.manual_seed(0)
    lin = MNFLinear(20, 7)
    perturb_mnf(lin, g)
    mnf_layer_case("mnf_linear_20x7", lin, torch.randn(33, 20, generator=g))

    torch.manual_seed(0)
    lin = MNFLinear(256, 128, n_flows_q=2, n_flows_r=1)
    perturb_mnf(lin, g)
    mnf_layer_case("mnf_linear_256x128", lin, torch.randn(48, 256, generator=g))

    torch.manual_seed(0)
    conv = MNFConv2d(2, 3, kernel_size=3)
    perturb_mnf(conv, g)
    mnf_layer_case("mnf_conv_2x3k3", conv, torch.rand(5, 2, 8, 8, generator=g))

    # MNF-LeNet, briefly trained on a synthetic 10-class template+noise problem
    torch.manual_seed(0)
    net = MNFLeNet()
    templates = torch.nn.functional.interpolate(
        torch.rand(10, 1, 7, 7, generator=g), size=28, mode="bilinear"
    )

    def batch(n):
        y = torch.randint(0, 10, (n,), generator=g)
        x = (templates[y] + 0.25 * torch.randn(n, 1, 28, 28, generator=g)).clamp(0, 1)
        return x, y

    opt = torch.optim.Adam(net.parameters())
    for _ in range(120):  # tests/test_mnf_mnist.py:28-43 loss
        xb, yb = batch(32)
        loss = torch.nn.functional.nll_loss(net(xb), yb) + 1e-3 * net.kl_div()
        opt.zero_grad()
        loss.backward()
        opt.step()
    xe, ye = batch(6)
    arrays = sd_np(net)
    arrays["x"] = xe.numpy()
    arrays["labels"] = ye.numpy()
    arrays["templates"] = templates.numpy()
    with torch.no_grad():
        t1 = NoiseTape()
        with record(t1):
            y = net(xe)
        arrays["fwd/y"] = y.numpy()
        arrays.update(t1.to_npz_dict("fwd_noise/"))
        t2 = NoiseTape()
        with record(t2):
            kl = net.kl_div()
        arrays["kl/value"] = kl.numpy()
        arrays.update(t2.to_npz_dict("kl_noise/"))
        acc = (net(batch(256)[0]).argmax(1) == batch(256)[1]).float().mean()
    print("lenet train acc (sanity, different labels -> chance):", float(acc), "loss", float(loss))
    save("mnf_lenet", **arrays)


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_flows()
    gen_mnf()
