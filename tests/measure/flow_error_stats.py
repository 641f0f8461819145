"""Error statistics of the CUDA flow kernels vs the fp32 oracle and an fp64 evaluation of it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
from oracle import flows_cpu
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden, random_flow_sd, t
from tests.test_flows_gpu import ORACLE_CASES


def stats(name, sd, specs, x):
    model = load_flow_model(specs, sd)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for inverse in (True, False):
        r32, l32 = flows_cpu.stack(sd, specs, x, inverse)
        r64, l64 = flows_cpu.stack(sd64, specs, x.double(), inverse)
        for kernel in (2, 3, "generic"):
            y, ld, _, _ = model._program().run(x.cuda(), inverse, kernel=kernel)
            y, ld = y.cpu().double(), ld.cpu().double()
            def q(e):
                return f"max {e.max():.2e} p99.9 {e.flatten().kthvalue(max(1, int(0.999 * e.numel()))).values:.2e}"
            ez, el = (y - r32[-1].double()).abs(), (ld - l32.double()).abs()
            ez64, el64 = (y - r64[-1]).abs(), (ld - l64).abs()
            oz, ol = (r32[-1].double() - r64[-1]).abs(), (l32.double() - l64).abs()
            fz = (ez <= 1e-5 * r64[-1].abs() + 1e-5).double().mean()
            fl = (el <= 1e-5 * l64.abs() + 1e-5).double().mean()
            print(f"{name} inv={int(inverse)} k={kernel}: z vs o32 [{q(ez)}] vs f64 [{q(ez64)}] | o32 vs f64 [{q(oz)}] frac {fz:.5f}"
                  f"\n      ld vs o32 [{q(el)}] vs f64 [{q(el64)}] | o32 vs f64 [{q(ol)}] frac {fl:.5f}", flush=True)


g = load_golden("nsfcl3_stack")
stats("golden_nsfcl3", golden_sd(g), golden_spec(g), t(g, "inv/x"))
gen = torch.Generator().manual_seed(11)
x = 1.5 * torch.randn(20001, 2, generator=gen)
specs = ORACLE_CASES["cfg2_shape"]
stats("seeded_cfg2", random_flow_sd(specs, seed=3, scale=0.6), specs, x)
