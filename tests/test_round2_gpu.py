"""Round-2 GPU parity additions (VERDICT r1 items 1, 9 and the ADVICE findings): the BASELINE config-5 shape on the
tensor-core path against the ORACLE (not against the repo's own fp32 kernel), standalone MaskedLinear / MADE calls,
the per-sample conv-z predict option, differentiable RNVP / sample_z entry points, cache hygiene after real calls and
CUDA-graph replays that draw fresh noise."""

import copy

import pytest
import torch

from oracle import mnf_cpu
from oracle.noise import NoiseTape
from tests.helpers import golden_sd, load_golden, t
from tests.test_mnf_gpu import _lenet, seeded_tape

pytestmark = pytest.mark.gpu


def _wide_layer(n=4096, seed=0):
    from torch_mnf.layers import MNFLinear

    torch.manual_seed(seed)
    layer = MNFLinear(n, n)
    with torch.no_grad():  # visible variance / q0 noise so that every term of the output matters
        layer.W_log_var += 5.0
        layer.q0_log_var += 8.0
        layer.W_mean *= 0.3
    return layer


def test_cfg5_shape_tensor_core_forward_vs_oracle():
    """BASELINE config 5: MNFLinear(4096, 4096), 64 input rows x 8 MC samples, identical injected noise, tcgen05 path
    against oracle.mnf_cpu.linear_forward at the stated tensor-core tolerance rtol 2e-3 (mnf_linear.py:46-64)."""
    n, rows, S = 4096, 64, 8
    layer = _wide_layer(n)
    sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(rows, n, generator=torch.Generator().manual_seed(1))
    R = rows * S
    draws = [("normal", (R, n)), ("bernoulli", (R, n)), ("bernoulli", (R, n)), ("normal", (R, n))]
    ref = mnf_cpu.linear_forward(sd, x.repeat(S, 1), seeded_tape(draws, 11))
    layer.cuda()
    layer.precision = "tf32"
    tape = seeded_tape(draws, 11)
    y = layer.forward_mc(x.cuda(), S, noise=tape).cpu()
    assert tape.pos == 4
    scale = float(ref.pow(2).mean().sqrt())
    torch.testing.assert_close(y, ref, rtol=2e-3, atol=2e-3 * scale)
    # the exact-fp32 path at the same shape (tighter class)
    layer.precision = "fp32"
    y32 = layer.forward_mc(x.cuda(), S, noise=seeded_tape(draws, 11)).cpu()
    torch.testing.assert_close(y32, ref, rtol=1e-4, atol=1e-4 * scale)


def test_cfg5_shape_kl_div_vs_oracle():
    """kl_div() of the config-5 layer with injected noise against oracle.mnf_cpu.linear_kl_div (mnf_linear.py:66-90)."""
    n = 4096
    layer = _wide_layer(n, seed=3)
    sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    draws = [("normal", (1, n)), ("bernoulli", (1, n)), ("bernoulli", (1, n)), ("normal", (n, n)),
             ("bernoulli", (1, n)), ("bernoulli", (1, n))]
    ref = mnf_cpu.linear_kl_div(sd, seeded_tape(draws, 5))
    layer.cuda()
    tape = seeded_tape(draws, 5)
    kl = layer.kl_div(noise=tape)
    assert tape.pos == len(draws)
    torch.testing.assert_close(kl.cpu(), ref, rtol=2e-5, atol=1e-3)
    terms = layer.__dict__["_last_kl_terms"].cpu()
    assert torch.isfinite(terms).all()


def test_mnf_linear_mc_wide_tensor_core_vs_oracle():
    """64 rows x 32 MC samples at 512 x 512 with injected noise: tensor-core pipeline (RNVP q-flow on tcgen05, variance
    GEMM once per distinct row, mean GEMM with noise epilogue) against the oracle."""
    from torch_mnf.layers import MNFLinear

    torch.manual_seed(0)
    layer = MNFLinear(512, 512)
    with torch.no_grad():
        layer.W_log_var += 6.0
        layer.q0_log_var += 8.0
    sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    x = torch.randn(64, 512, generator=torch.Generator().manual_seed(2))
    R = 64 * 32
    draws = [("normal", (R, 512)), ("bernoulli", (R, 512)), ("bernoulli", (R, 512)), ("normal", (R, 512))]
    ref = mnf_cpu.linear_forward(sd, x.repeat(32, 1), seeded_tape(draws, 5))
    layer.cuda()
    layer.precision = "tf32"
    out = layer.forward_mc(x.cuda(), 32, noise=seeded_tape(draws, 5)).cpu()
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3 * float(ref.pow(2).mean().sqrt()))


def test_masked_linear_and_made_forward_standalone():
    """made.py:22-23 / :43-57: MaskedLinear(x) = x @ (W.T * mask) + b and MADE(x) as a module call, against torch fp32
    on the CPU with the same parameters and the reference's masks; gradients against torch autograd."""
    from torch_mnf.layers import MADE, MaskedLinear

    torch.manual_seed(0)
    made = MADE(8, [16, 16], 16, natural_ordering=True)
    x = torch.randn(300, 8)

    def ref_fn(xc, params):
        h = xc
        lins = [m for m in made if isinstance(m, MaskedLinear)]
        for i, m in enumerate(lins):
            w, b = params[2 * i], params[2 * i + 1]
            h = h @ (w.T * m.mask.cpu().float()) + b
            if i + 1 < len(lins):
                h = torch.relu(h)
        return h

    cpu_params = [p.detach().clone().requires_grad_(True) for p in made.parameters()]
    xc = x.clone().requires_grad_(True)
    with torch.enable_grad():
        ref = ref_fn(xc, cpu_params)
        ref.square().sum().backward()
    made.cuda()
    with torch.no_grad():
        y = made(x.cuda())
    torch.testing.assert_close(y.cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    # autoregressive property: output i (both halves) does not depend on inputs >= i
    x2 = x.clone()
    x2[:, 5:] += 1.0
    with torch.no_grad():
        y2 = made(x2.cuda()).cpu()
    assert torch.equal(y2[:, :6], y.cpu()[:, :6]) and torch.equal(y2[:, 8:14], y.cpu()[:, 8:14])
    with torch.enable_grad():
        xg = x.cuda().requires_grad_(True)
        out = made(xg)
        out.square().sum().backward()
    torch.testing.assert_close(xg.grad.cpu(), xc.grad, rtol=1e-4, atol=1e-4)
    for p, q in zip(made.parameters(), cpu_params):
        torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=1e-4, atol=1e-4 * float(q.grad.abs().max()) + 1e-6)
    lin = made[0]
    with torch.no_grad():
        one = lin(x.cuda()[:, None, :])  # leading dims are kept
    assert one.shape == (300, 1, 16)


def _per_sample_draws(S, R):
    return [
        ("normal", (S, 20)), ("bernoulli", (S, 20)), ("bernoulli", (S, 20)), ("normal", (R, 20, 24, 24)),
        ("normal", (S, 50)), ("bernoulli", (S, 50)), ("bernoulli", (S, 50)), ("normal", (R, 50, 8, 8)),
        ("normal", (R, 800)), ("bernoulli", (R, 800)), ("bernoulli", (R, 800)), ("normal", (R, 50)),
        ("normal", (R, 50)), ("bernoulli", (R, 50)), ("bernoulli", (R, 50)), ("normal", (R, 10)),
    ]


@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_lenet_per_sample_conv_z_matches_separate_reference_calls(precision):
    """with per_sample_conv_z every MC sample has its own conv z -- sample s must equal ONE oracle call
    (mnf_lenet.py:13-26 over mnf_conv.py:80-88) fed sample s's rows of every draw."""
    g, net = _lenet()
    net.precision = precision
    sd = golden_sd(g)
    gen = torch.Generator().manual_seed(4)
    labels = torch.randint(0, 10, (6,), generator=gen)
    x = (t(g, "templates")[labels] + 0.25 * torch.randn(6, 1, 28, 28, generator=gen)).clamp(0, 1)
    B, S = 6, 5
    R = B * S
    big = seeded_tape(_per_sample_draws(S, R), 9)
    y = net(x.cuda(), noise=seeded_tape(_per_sample_draws(S, R), 9), n_samples=S, per_sample_conv_z=True).cpu()
    assert y.shape == (R, 10)
    refs = []
    for s in range(S):
        rows = slice(s * B, (s + 1) * B)
        draws = []
        for i, (kind, ten) in enumerate(big.draws):
            if i in (0, 4):
                draws.append((kind, ten[s].clone()))  # randn_like[n_out]
            elif i in (1, 2, 5, 6):
                draws.append((kind, ten[s:s + 1].clone()))  # bernoulli[1, n_out]
            else:
                draws.append((kind, ten[rows].clone()))
        refs.append(mnf_cpu.lenet_forward(sd, x, NoiseTape(draws)))
    ref = torch.cat(refs)
    if precision == "fp32":
        torch.testing.assert_close(y, ref, rtol=1e-4, atol=2e-4)
    else:
        torch.testing.assert_close(y.exp(), ref.exp(), rtol=2e-2, atol=2e-3)
    assert torch.equal(y.exp().view(S, B, 10).mean(0).argmax(1), ref.exp().view(S, B, 10).mean(0).argmax(1))
    # shared-z mode differs (one z for all samples) -- the option is not a no-op
    shared = net(x.cuda(), n_samples=S, seed=3).cpu()
    per = net(x.cuda(), n_samples=S, seed=3, per_sample_conv_z=True).cpu()
    assert not torch.allclose(shared, per)
    # sharding invariance in Philox mode: two halves of the samples reproduce the full call
    S2 = 4
    full = net(x.cuda(), n_samples=S2, seed=21, per_sample_conv_z=True)
    lo = net(x.cuda(), n_samples=S2 // 2, seed=21, per_sample_conv_z=True)
    hi = net(x.cuda(), n_samples=S2 // 2, seed=21, per_sample_conv_z=True, row_offset=(S2 // 2) * B)
    torch.testing.assert_close(torch.cat([lo, hi]), full, rtol=1e-5, atol=1e-5)


def test_rnvp_and_sample_z_are_differentiable_like_the_reference():
    """rnvp.py:25-39 / mnf_linear.py:58-64 are differentiable under torch autograd; the drop-in entry points must not
    silently return graph-less tensors when grad is enabled."""
    import torch_mnf.flows as nf
    from torch_mnf.layers import MNFConv2d, MNFLinear

    torch.manual_seed(0)
    with torch.enable_grad():
        f = nf.RNVP(10, h_sizes=(12,)).cuda()
        z = torch.randn(7, 10, device="cuda", requires_grad=True)
        x, ld = f.forward(z)
        assert x.grad_fn is not None and ld.grad_fn is not None
        (x.sum() + ld.sum()).backward()
        assert z.grad is not None and f.t.weight.grad is not None
        stack = nf.NormalizingFlow([nf.RNVP(10, h_sizes=(12,)) for _ in range(2)]).cuda()
        xs, ld = stack.forward(z.detach())
        assert len(xs) == 3 and xs[-1].grad_fn is not None
        lin = MNFLinear(10, 4).cuda()
        zz, ldq = lin.sample_z(5)
        assert zz.shape == (5, 10) and zz.grad_fn is not None and ldq.shape == (5,)
        zz.sum().backward()
        assert lin.q0_mean.grad is not None
        conv = MNFConv2d(2, 6, 3).cuda()
        zc, ldc = conv.sample_z()
        assert zc.shape == (1, 6) and zc.grad_fn is not None
    with torch.no_grad():
        x2, _ = f.forward(z.detach())
        assert x2.grad_fn is None


def test_modules_copy_and_pickle_after_real_calls():
    import io

    from tests.helpers import golden_spec, load_flow_model

    g, net = _lenet()
    x = t(g, "x").cuda()
    with torch.no_grad():
        net(x)
        net.kl_div()
    c = copy.deepcopy(net)
    buf = io.BytesIO()
    torch.save(net, buf)
    with torch.no_grad():
        a = c(x, seed=5)
        b = net(x, seed=5)
    assert torch.equal(a, b)
    gm = load_golden("maf9_d64")
    model = load_flow_model(golden_spec(gm), golden_sd(gm))
    xm = t(gm, "inv/x").cuda()
    with torch.no_grad():
        z0, ld0 = model.inverse(xm)
        model.flows[0].inverse(xm)
    m2 = copy.deepcopy(model)
    buf = io.BytesIO()
    torch.save(model, buf)
    with torch.no_grad():
        z1, ld1 = m2.inverse(xm)
    assert torch.equal(z0[-1], z1[-1]) and torch.equal(ld0, ld1)


def test_packed_parameters_follow_every_kind_of_update():
    """Evaluate, change parameters in ways that do not bump a version counter, evaluate again."""
    from tests.helpers import golden_spec, load_flow_model

    g = load_golden("nsfcl3_stack")
    model = load_flow_model(golden_spec(g), golden_sd(g), return_intermediates=False)
    x = t(g, "inv/x").cuda()
    with torch.no_grad():
        lp0 = model.log_prob(x).clone()
        mid = model.flows[5].f1[2].weight  # a tensor in the middle of the packed blob
        mid.data = mid.data * 1.5
        lp1 = model.log_prob(x).clone()
        assert not torch.allclose(lp0, lp1)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        sd["flows.5.f1.2.weight"] = sd["flows.5.f1.2.weight"] / 1.5
        model.load_state_dict(sd, assign=True)
        lp2 = model.log_prob(x)
        torch.testing.assert_close(lp2, lp0, rtol=1e-5, atol=2e-5)  # w * 1.5 / 1.5 is not bit-exact


def test_graphed_training_replays_invalidate_packed_blobs():
    """graphed_training_step replays update parameters without bumping versions: an evaluation between replays must see
    the new weights."""
    import torch_mnf.flows as nf
    from torch.distributions import MultivariateNormal
    from torch_mnf.graphs import graphed_training_step

    torch.manual_seed(0)
    with torch.enable_grad():
        model = nf.NormalizingFlowModel(MultivariateNormal(torch.zeros(2), torch.eye(2)),
                                        [nf.AffineHalfFlow(2, parity=bool(i % 2), h_sizes=(8, 8)) for i in range(2)]).cuda()
        x = torch.randn(256, 2, device="cuda")
        opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)
        step = graphed_training_step(model, lambda m, xb: -m.log_prob(xb).mean(), opt, (x,))
        with torch.no_grad():
            before = model.log_prob(x).mean().item()
        for _ in range(20):
            step(x)
        with torch.no_grad():
            after = model.log_prob(x).mean().item()
        ref = nf.NormalizingFlowModel(MultivariateNormal(torch.zeros(2), torch.eye(2)),
                                      [nf.AffineHalfFlow(2, parity=bool(i % 2), h_sizes=(8, 8)) for i in range(2)]).cuda()
        ref.load_state_dict(model.state_dict())
        with torch.no_grad():
            fresh = ref.log_prob(x).mean().item()
    assert after > before + 1e-3, (before, after)
    assert abs(after - fresh) < 1e-5, (after, fresh)


def test_graphed_inference_draws_fresh_noise_per_replay():
    from torch_mnf.graphs import graphed_inference
    from torch_mnf.layers import MNFLinear

    torch.manual_seed(0)
    layer = MNFLinear(32, 8).cuda()
    with torch.no_grad():
        layer.W_log_var += 6.0
    x = torch.randn(16, 32, device="cuda")
    call = graphed_inference(lambda xb: layer(xb), (x,))
    a = call(x).clone()
    b = call(x).clone()
    assert not torch.equal(a, b)  # independent MC samples, not a frozen seed
    m = torch.stack([call(x).clone() for _ in range(200)])
    with torch.no_grad():
        eager = torch.stack([layer(x) for _ in range(200)])
    # same distribution as eager calls: means agree within 6 standard errors of the 200-sample means
    se = (eager.std(0) + m.std(0)) / 200**0.5
    assert ((m.mean(0) - eager.mean(0)).abs() < 6 * se + 1e-3).all()


def test_lenet_kl_div_multi_matches_loop():
    """MNFLeNet.kl_div through mnf_kl_div_fused_multi (four layers, three launches) against the per-layer loop under
    the same seeds: same draws, same value, same per-layer terms; deep copies and training mode keep working."""
    import copy

    from torch_mnf import _lib
    from torch_mnf.models import MNFLeNet

    torch.manual_seed(0)
    m = MNFLeNet().cuda()
    layers = [layer for layer in m if hasattr(layer, "kl_div")]
    with torch.no_grad():
        torch.manual_seed(5)
        ref_terms = []
        for layer in layers:
            layer.kl_div()
            ref_terms.append(layer.__dict__["_last_kl_terms"].clone())
        ref = sum(t[0] for t in ref_terms)
        torch.manual_seed(5)
        _lib.launch_stats(reset=True)
        got = m.kl_div()
        sites = _lib.launch_stats()
        assert sites == {"kl_flows_multi_kernel": 1, "kl_rows_multi_kernel": 1, "kl_final_multi_kernel": 1}, sites
        torch.testing.assert_close(got, ref, rtol=1e-6, atol=0)
        for layer, t in zip(layers, ref_terms):
            torch.testing.assert_close(layer.__dict__["_last_kl_terms"], t, rtol=1e-6, atol=1e-6)
        torch.manual_seed(5)
        torch.testing.assert_close(copy.deepcopy(m).kl_div(), ref, rtol=1e-6, atol=0)
    with torch.enable_grad():  # (this module's tests run with grad disabled: tests/conftest.py)
        kl = m.kl_div()  # trainable parameters and grad enabled: the differentiable path
    assert kl.requires_grad
