"""Gradients of the MNF layers' training path (csrc/mnf_train.cu through torch_mnf/layers/_train.py) against torch
autograd over the fp64 CPU oracle with the SAME injected noise, and an end-to-end training run with the reference's
loss (nll + kl_div * 1e-3, tests/test_mnf_mnist.py:28-31)."""

import pytest
import torch

from oracle import mnf_cpu
from tests.helpers import golden_sd, golden_tape, load_golden, t

pytestmark = pytest.mark.gpu


def _sd64(sd):
    return {k: (v.double().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}


def _compare(layer, sd64, what):
    checked = 0
    for k, p in layer.named_parameters():
        ref = sd64[k].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{what}: unexpected gradient on {k}"
            continue
        assert p.grad is not None, f"{what}: no gradient reached {k}"
        scale = float(ref.abs().max()) + 1e-9
        err = float((p.grad.cpu().double() - ref).abs().max())
        assert err <= 2e-4 * scale + 1e-7, f"{what} {k}: max err {err:.3e} vs scale {scale:.3e}"
        checked += 1
    assert checked >= 8


@pytest.mark.parametrize("name,n_in,n_out,kw", [("mnf_linear_20x7", 20, 7, {}),
                                                ("mnf_linear_256x128", 256, 128, {"n_flows_r": 1})])
def test_mnf_linear_forward_gradients(name, n_in, n_out, kw):
    from torch_mnf.layers import MNFLinear

    g = load_golden(name)
    sd = golden_sd(g)
    layer = MNFLinear(n_in, n_out, **kw)
    layer.load_state_dict(sd, strict=True)
    layer.cuda()
    x = t(g, "x")
    w = torch.randn(x.size(0), n_out, generator=torch.Generator().manual_seed(1), dtype=torch.float64)

    sd64 = _sd64(sd)
    x64 = x.double().requires_grad_()
    y_ref = mnf_cpu.linear_forward(sd64, x64, golden_tape(g, "fwd_noise/"))
    (y_ref * w).sum().backward()

    xg = x.cuda().requires_grad_()
    y = layer(xg, noise=golden_tape(g, "fwd_noise/"))
    assert y.requires_grad
    torch.testing.assert_close(y.detach().cpu(), t(g, "fwd/y"), rtol=1e-4, atol=1e-5)
    (y * w.cuda().float()).sum().backward()
    scale = float(x64.grad.abs().max())
    assert float((xg.grad.cpu().double() - x64.grad).abs().max()) <= 2e-4 * scale
    _compare(layer, sd64, name)


@pytest.mark.parametrize("name,n_in,n_out,kw", [("mnf_linear_20x7", 20, 7, {}),
                                                ("mnf_linear_256x128", 256, 128, {"n_flows_r": 1})])
def test_mnf_linear_kl_div_gradients(name, n_in, n_out, kw):
    from torch_mnf.layers import MNFLinear

    g = load_golden(name)
    sd = golden_sd(g)
    layer = MNFLinear(n_in, n_out, **kw)
    layer.load_state_dict(sd, strict=True)
    layer.cuda()
    sd64 = _sd64(sd)
    mnf_cpu.linear_kl_div(sd64, golden_tape(g, "kl_noise/")).backward()
    kl = layer.kl_div(noise=golden_tape(g, "kl_noise/"))
    assert kl.requires_grad and kl.shape == ()
    torch.testing.assert_close(kl.detach().cpu(), t(g, "kl/value"), rtol=2e-5, atol=1e-4)
    kl.backward()
    _compare(layer, sd64, name + " kl_div")


def test_mnf_conv_gradients():
    from torch_mnf.layers import MNFConv2d

    g = load_golden("mnf_conv_2x3k3")
    sd = golden_sd(g)
    layer = MNFConv2d(2, 3, kernel_size=3)
    layer.load_state_dict(sd, strict=True)
    layer.cuda()
    x = t(g, "x")
    y_gold = t(g, "fwd/y")
    w = torch.randn(y_gold.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)

    sd64 = _sd64(sd)
    x64 = x.double().requires_grad_()
    y_ref = mnf_cpu.conv_forward(sd64, x64, golden_tape(g, "fwd_noise/"))
    ((y_ref * w).sum() + mnf_cpu.conv_kl_div(sd64, golden_tape(g, "kl_noise/"))).backward()

    xg = x.cuda().requires_grad_()
    y = layer(xg, noise=golden_tape(g, "fwd_noise/"))
    torch.testing.assert_close(y.detach().cpu(), y_gold, rtol=1e-4, atol=1e-5)
    kl = layer.kl_div(noise=golden_tape(g, "kl_noise/"))
    torch.testing.assert_close(kl.detach().cpu(), t(g, "kl/value"), rtol=2e-5, atol=1e-4)
    ((y * w.cuda().float()).sum() + kl).backward()
    scale = float(x64.grad.abs().max())
    assert float((xg.grad.cpu().double() - x64.grad).abs().max()) <= 2e-4 * scale
    _compare(layer, sd64, "mnf_conv_2x3k3")


def test_mnf_lenet_gradients():
    """The reference's training loss (tests/test_mnf_mnist.py:28-31) on the LeNet fixture: every parameter gradient
    against the fp64 oracle under the same noise."""
    from torch_mnf.models import MNFLeNet

    g = load_golden("mnf_lenet")
    sd = golden_sd(g)
    net = MNFLeNet()
    net.load_state_dict(sd, strict=True)
    net.cuda()
    x = t(g, "x")
    labels = torch.arange(x.size(0)) % 10

    sd64 = _sd64(sd)
    ref = torch.nn.functional.nll_loss(mnf_cpu.lenet_forward(sd64, x.double(), golden_tape(g, "fwd_noise/")), labels)
    ref = ref + 1e-3 * mnf_cpu.lenet_kl_div(sd64, golden_tape(g, "kl_noise/"))
    ref.backward()

    logp = net(x.cuda(), noise=golden_tape(g, "fwd_noise/"))
    torch.testing.assert_close(logp.detach().cpu(), t(g, "fwd/y"), rtol=1e-4, atol=2e-5)
    loss = torch.nn.functional.nll_loss(logp, labels.cuda()) + 1e-3 * net.kl_div(noise=golden_tape(g, "kl_noise/"))
    torch.testing.assert_close(loss.detach().cpu().double(), ref.detach(), rtol=1e-5, atol=1e-5)
    loss.backward()
    _compare(net, sd64, "mnf_lenet")


def test_mnf_lenet_trains():
    """tests/test_mnf_mnist.py:34-58 on synthetic digits (no dataset offline): noisy copies of 10 fixed templates,
    batch 32, Adam, loss = nll + 1e-3 * kl_div, stop once a batch reaches 95 %; validation accuracy must exceed 0.8."""
    from torch_mnf.models import MNFLeNet

    torch.manual_seed(0)
    templates = t(load_golden("mnf_lenet"), "templates").cuda()

    def batch(n):
        y = torch.randint(0, 10, (n,), device="cuda")
        return (templates[y] + 0.25 * torch.randn(n, 1, 28, 28, device="cuda")).clamp(0, 1), y

    net = MNFLeNet().cuda()
    adam = torch.optim.Adam(net.parameters())
    for _step in range(400):
        x, y = batch(32)
        adam.zero_grad()
        preds = net(x)
        loss = torch.nn.functional.nll_loss(preds, y) + 1e-3 * net.kl_div()
        loss.backward()
        adam.step()
        if float((preds.argmax(1) == y).float().mean()) > 0.95:
            break
    x_val, y_val = batch(500)
    with torch.no_grad():
        val_acc = float((net(x_val).argmax(1) == y_val).float().mean())
    assert val_acc > 0.8, f"val_acc {val_acc:.3f} after {_step + 1} steps"


def test_graphed_training_step():
    """The whole MNF-LeNet training step (forward, kl_div, backward, Adam) captured in one CUDA graph and replayed."""
    from torch_mnf.graphs import graphed_training_step
    from torch_mnf.models import MNFLeNet

    torch.manual_seed(0)
    templates = t(load_golden("mnf_lenet"), "templates").cuda()

    def batch(n):
        y = torch.randint(0, 10, (n,), device="cuda")
        return (templates[y] + 0.25 * torch.randn(n, 1, 28, 28, device="cuda")).clamp(0, 1), y

    net = MNFLeNet().cuda()
    adam = torch.optim.Adam(net.parameters(), capturable=True)

    def loss_fn(model, x, y):
        return torch.nn.functional.nll_loss(model(x), y) + 1e-3 * model.kl_div()

    step = graphed_training_step(net, loss_fn, adam, batch(32))
    losses = [float(step(*batch(32))) for _ in range(150)]
    assert all(l == l for l in losses) and sum(losses[-10:]) < sum(losses[:10])
    x_val, y_val = batch(500)
    with torch.no_grad():
        val_acc = float((net(x_val).argmax(1) == y_val).float().mean())
    assert val_acc > 0.8, f"val_acc {val_acc:.3f}"


def test_gemm_f32_all_transpositions():
    from torch_mnf.layers import _train

    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in ((1, 7, 20), (33, 70, 130), (128, 64, 64), (5, 300, 1)):
        A = torch.randn(M, K, device="cuda", generator=g)
        B = torch.randn(K, N, device="cuda", generator=g)
        bias = torch.randn(N, device="cuda", generator=g)
        C0 = torch.randn(M, N, device="cuda", generator=g)
        ref = (A.double() @ B.double() + bias.double() + 0.5 * C0.double()).float()
        for ta in (False, True):
            for tb in (False, True):
                out = C0.clone()
                _train.gemm(A.t().contiguous() if ta else A, B.t().contiguous() if tb else B, ta, tb, bias=bias,
                            out=out, beta=0.5)
                torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-4)


def test_mnf_feed_forward_trains():
    """Three Gaussian blobs in 16-d, MNFFeedForward([16, 32, 3]), loss = nll + 1e-3 * kl_div, Adam: the loss must fall
    and the classifier must fit the training points."""
    from torch_mnf.models import MNFFeedForward

    torch.manual_seed(0)
    centers = 3.0 * torch.randn(3, 16)
    y = torch.arange(3).repeat_interleave(64)
    x = (centers[y] + torch.randn(192, 16)).cuda()
    y = y.cuda()
    model = MNFFeedForward([16, 32, 3]).cuda()
    adam = torch.optim.Adam(model.parameters(), lr=1e-2)
    losses = []
    for _ in range(60):
        adam.zero_grad()
        logp = torch.log_softmax(model(x), dim=-1)
        loss = torch.nn.functional.nll_loss(logp, y) + 1e-3 * model.kl_div()
        loss.backward()
        adam.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]
    model.eval()
    with torch.no_grad():
        acc = float((model(x).argmax(1) == y).float().mean())
    assert acc > 0.9, f"accuracy {acc}"
