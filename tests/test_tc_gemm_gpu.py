"""Tensor-core (tcgen05 / TF32 / TMA) path: plain linear and the MNFLinear forward, against an fp32
reference with the rtol 2e-3 BASELINE.json states for tensor-core GEMM outputs (the absolute floor is
2e-3 x the row's typical magnitude: TF32 rounds each product to 10 mantissa bits)."""

import pytest
import torch

from tests.helpers import golden_sd, golden_tape, load_golden, t

pytestmark = pytest.mark.gpu


def tc_linear(A, W, bias=None, relu=False):
    import torch_mnf.layers  # noqa: F401
    from torch_mnf import _lib

    out = torch.empty(A.size(0), W.size(0), device=A.device)
    rc = _lib.lib().mnf_tc_linear(A.data_ptr(), W.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(),
                                  A.size(0), W.size(0), A.size(1), int(relu), 0, _lib.stream_ptr(A.device))
    _lib.check(rc, "mnf_tc_linear")
    return out


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (256, 256, 64), (1000, 320, 132), (77, 8, 36), (4096, 512, 1024),
                                   (130, 1000, 4096)])
def test_tc_linear_matches_fp32(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K**0.5
    b = torch.randn(N, device="cuda", generator=g)
    pre = A.double() @ W.double().T
    # the raw entry point feeds unrounded fp32: the tensor core truncates both operands to TF32 (up to 2^-10
    # each, biased), so the bound is 4e-3 of the typical |a.w|; the MNF path pre-rounds to nearest (2e-3 below)
    atol = 4e-3 * float(pre.pow(2).mean().sqrt())
    ref = (pre + b.double()).float()
    out = tc_linear(A, W, b)
    torch.testing.assert_close(out, ref, rtol=4e-3, atol=atol)
    out = tc_linear(A, W, None, relu=True)
    torch.testing.assert_close(out, torch.relu(pre).float(), rtol=4e-3, atol=atol)


def test_mnf_linear_tf32_vs_golden():
    from torch_mnf.layers import MNFLinear

    g = load_golden("mnf_linear_256x128")
    layer = MNFLinear(256, 128, n_flows_r=1)
    layer.load_state_dict(golden_sd(g), strict=True)
    layer.cuda()
    layer.precision = "tf32"
    y = layer(t(g, "x").cuda(), noise=golden_tape(g, "fwd_noise/"))
    ref = t(g, "fwd/y")
    torch.testing.assert_close(y.cpu(), ref, rtol=2e-3, atol=2e-3 * float(ref.pow(2).mean().sqrt()))
    layer.precision = "fp32"
    y32 = layer(t(g, "x").cuda(), noise=golden_tape(g, "fwd_noise/"))
    torch.testing.assert_close(y32.cpu(), ref, rtol=1e-4, atol=1e-5 * float(ref.abs().mean()))


def test_mnf_linear_tf32_mc_replication_wide():
    """Config-5 shape in small: 64 input rows x 32 MC samples, n_in = n_out = 512, Philox noise; the tensor-core
    result must match the fp32 SIMT path (same Philox draws) to 2e-3."""
    from torch_mnf.layers import MNFLinear
    from torch_mnf.layers._mnf_ops import Noise

    torch.manual_seed(0)
    layer = MNFLinear(512, 512).cuda()
    with torch.no_grad():
        layer.W_log_var += 6.0
        layer.q0_log_var += 8.0
    x = torch.randn(64, 512, device="cuda")
    layer.precision = "fp32"
    ref = layer.forward_mc(x, 32, noise=Noise(None, x.device, 0, seed=5))
    layer.precision = "tf32"
    out = layer.forward_mc(x, 32, noise=Noise(None, x.device, 0, seed=5))
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3 * float(ref.pow(2).mean().sqrt()))


def test_maf_stack_tensor_core_density_vs_golden():
    """BASELINE config 3 shape: MAF x9, D = 64, density direction on the tensor cores (tf32 tolerance class)."""
    from tests.helpers import golden_spec, load_flow_model

    g = load_golden("maf9_d64")
    model = load_flow_model(golden_spec(g), golden_sd(g))
    for f in model.flows:
        f.precision = "tf32"
    x = t(g, "inv/x").cuda()
    zs, ld = model.inverse(x)
    assert len(zs) == 10
    ref_z, ref_ld = t(g, "inv/z"), t(g, "inv/ld")
    torch.testing.assert_close(zs[-1].cpu(), ref_z, rtol=2e-3, atol=2e-3 * float(ref_z.pow(2).mean().sqrt()))
    torch.testing.assert_close(ld.cpu(), ref_ld, rtol=2e-3, atol=2e-3 * float(ref_ld.pow(2).mean().sqrt()) + 2e-3)
    torch.testing.assert_close(zs[5].cpu(), t(g, "inv/z_mid"), rtol=2e-3, atol=2e-3 * float(ref_z.pow(2).mean().sqrt()))
    # single flow through the module API, and the exact-fp32 path for comparison
    z1, ld1 = model.flows[8].inverse(x)
    torch.testing.assert_close(z1, zs[1], rtol=1e-6, atol=1e-6)
    for f in model.flows:
        f.precision = "fp32"
    zs32, ld32 = model.inverse(x)
    torch.testing.assert_close(zs32[-1].cpu(), ref_z, rtol=1e-5, atol=2e-5)
    model.return_intermediates = False
    for f in model.flows:
        f.precision = "tf32"
    zs2, ld2 = model.inverse(x)
    assert len(zs2) == 2 and torch.equal(zs2[-1], zs[-1])
    torch.testing.assert_close(ld2, ld, rtol=1e-5, atol=1e-5)  # row sums are accumulated with atomics


def test_mnf_lenet_mc_pipeline_tensor_cores():
    """MNF-LeNet Monte-Carlo pipeline (conv1 moments once per image, conv2 / fc1 on tensor cores) vs the oracle on
    identical injected noise: tf32 tolerance class on the log-probabilities, exact MC-mean argmax."""
    from oracle import mnf_cpu
    from tests.test_mnf_gpu import _lenet, lenet_draws, seeded_tape

    g, net = _lenet()
    sd = golden_sd(g)
    gen = torch.Generator().manual_seed(7)
    labels = torch.randint(0, 10, (32,), generator=gen)
    x = (t(g, "templates")[labels] + 0.25 * torch.randn(32, 1, 28, 28, generator=gen)).clamp(0, 1)
    S, R = 16, 512
    assert R >= net.TC_MIN_ROWS
    ref = mnf_cpu.lenet_forward(sd, x.repeat(S, 1, 1, 1), seeded_tape(lenet_draws(R), 3))
    y = net(x.cuda(), noise=seeded_tape(lenet_draws(R), 3), n_samples=S).cpu()
    # log-probs of confidently rejected classes are very negative; compare on the probability scale as well
    torch.testing.assert_close(y.exp(), ref.exp(), rtol=2e-2, atol=2e-3)
    big = ref > -10
    torch.testing.assert_close(y[big], ref[big], rtol=2e-3, atol=2e-2)
    p_ref, p = ref.exp().view(S, 32, 10).mean(0), y.exp().view(S, 32, 10).mean(0)
    assert torch.equal(p.argmax(1), p_ref.argmax(1))
    net.precision = "fp32"
    y32 = net(x.cuda(), noise=seeded_tape(lenet_draws(R), 3), n_samples=S).cpu()
    torch.testing.assert_close(y32, ref, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("n_imgs", [1, 7, 300])
def test_implicit_gemm_conv_matches_exact_path(n_imgs):
    """MNF-LeNet's second conv (20 -> 50 channels, 5x5 on 12x12) through the implicit-GEMM tensor-core kernel (operand tiles
    generated in shared memory, odd image counts leave a half-filled last tile) against the exact-fp32 kernel with the
    same injected noise; tf32 tolerance class."""
    from tests.test_mnf_gpu import seeded_tape
    from torch_mnf import _lib
    from torch_mnf.layers import MNFConv2d
    from torch_mnf.layers import _mnf_ops as ops

    torch.manual_seed(1)
    layer = MNFConv2d(20, 50, kernel_size=5).cuda()
    with torch.no_grad():
        layer.W_log_var.add_(6.0)  # visible variance term
    assert _lib.lib().mnf_conv_tc_workspace(n_imgs, 20, 12, 12, 50, 5) < 3 * n_imgs * 64 * 64 + 2 * 64 * 512 + 200
    x = torch.rand(n_imgs, 20, 12, 12, generator=torch.Generator().manual_seed(2)).cuda()
    draws = [("normal", (50,)), ("bernoulli", (1, 50)), ("bernoulli", (1, 50)), ("normal", (n_imgs, 50, 8, 8))]
    ref = layer(x, noise=seeded_tape(draws, 5), relu_pool=True)
    noise = ops.Noise(seeded_tape(draws, 5), x.device)
    z, _ = layer.sample_z(noise)
    got = ops.conv_forward_tc(layer, x, z, noise)
    assert got.shape == ref.shape == (n_imgs, 50, 4, 4)
    torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3 * float(ref.abs().mean()))
