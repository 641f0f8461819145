"""Training-step timings: forward + backward + Adam of (a) the reference's flow regression stacks
(tests/test_flows.py) and (b) MNF-LeNet with loss = nll + 1e-3 kl_div (tests/test_mnf_mnist.py), on the CUDA path and,
for comparison, on the CPU oracle port with torch autograd (same arithmetic as the reference) on the host cores."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch
from torch.distributions import MultivariateNormal

import torch_mnf.flows as nf
from oracle import flows_cpu, mnf_cpu
from oracle.noise import FreshNoise
from tests.helpers import golden_sd, load_golden, t
from torch_mnf import data
from torch_mnf.models import MNFLeNet

out = {}


def gpu_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_ms(fn, iters=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e3


# (a) flows: ActNorm + NSF_CL x2 (test_nsfcl_with_actnorm), batch 128 and 2^16
specs = [{"type": "ActNormFlow", "dim": 2, "scale": True, "shift": True}, {"type": "NSF_CL", "dim": 2, "K": 8, "B": 3, "n_h": 16}] * 2
for n in (128, 1 << 16):
    torch.manual_seed(0)
    flows = [nf.ActNormFlow(dim=2), nf.NSF_CL(dim=2, K=8, B=3, n_h=16), nf.ActNormFlow(dim=2), nf.NSF_CL(dim=2, K=8, B=3, n_h=16)]
    model = nf.NormalizingFlowModel(MultivariateNormal(torch.zeros(2), torch.eye(2)), flows).cuda()
    x = data.sample_moons(n).cuda()
    adam = torch.optim.Adam(model.parameters())

    def step():
        _, ld = model.inverse(x)
        loss = -(ld + model.base_log_prob(x)).sum()
        model.zero_grad()
        loss.backward()
        adam.step()

    g = gpu_ms(step)
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    adam_c = torch.optim.Adam(params)
    xc = x.cpu()

    def cstep():
        loss = -flows_cpu.log_prob(sd, specs, xc).sum()
        adam_c.zero_grad()
        loss.backward()
        adam_c.step()

    c = cpu_ms(cstep, iters=3 if n > 1000 else 10)
    out[f"flow_nsfcl_actnorm_train_step_b{n}"] = {"gpu_ms": g, "cpu_port_ms": c, "cpu_threads": torch.get_num_threads()}
    print(f"flow training step, batch {n}: GPU {g:.3f} ms, CPU oracle+autograd {c:.1f} ms", flush=True)

# (b) MNF-LeNet, batch 32
torch.manual_seed(0)
templates = t(load_golden("mnf_lenet"), "templates")
y = torch.randint(0, 10, (32,))
xb = (templates[y] + 0.25 * torch.randn(32, 1, 28, 28)).clamp(0, 1)
net = MNFLeNet().cuda()
adam = torch.optim.Adam(net.parameters())
xg, yg = xb.cuda(), y.cuda()


def lstep():
    adam.zero_grad()
    loss = torch.nn.functional.nll_loss(net(xg), yg) + 1e-3 * net.kl_div()
    loss.backward()
    adam.step()


g = gpu_ms(lstep)
sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
adam_c = torch.optim.Adam([v for v in sd.values() if v.requires_grad])


def lcstep():
    adam_c.zero_grad()
    loss = torch.nn.functional.nll_loss(mnf_cpu.lenet_forward(sd, xb, FreshNoise()), y) + 1e-3 * mnf_cpu.lenet_kl_div(sd, FreshNoise())
    loss.backward()
    adam_c.step()


c = cpu_ms(lcstep, iters=5)
# the same step captured in a CUDA graph (torch_mnf.graphs)
from torch_mnf.graphs import graphed_training_step

net2 = MNFLeNet().cuda()
adam2 = torch.optim.Adam(net2.parameters(), capturable=True)
gstep = graphed_training_step(net2, lambda m, a, b: torch.nn.functional.nll_loss(m(a), b) + 1e-3 * m.kl_div(), adam2, (xg, yg))
gg = gpu_ms(lambda: gstep(xg, yg))
out["mnf_lenet_train_step_b32"] = {"gpu_ms": g, "gpu_graph_ms": gg, "cpu_port_ms": c, "cpu_threads": torch.get_num_threads()}
print(f"MNF-LeNet training step, batch 32: GPU {g:.3f} ms eager, {gg:.3f} ms as one CUDA graph, CPU oracle+autograd {c:.1f} ms", flush=True)
print(json.dumps(out))
