"""cfg3: MAF x9, D=64, density evaluation; tensor-core chain vs exact-fp32 interpreter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden

g = load_golden("maf9_d64")
model = load_flow_model(golden_spec(g), golden_sd(g), return_intermediates=False)


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)


for prec, n in (("tf32", 1 << 20), ("fp32", 1 << 20)):
    for f in model.flows:
        f.precision = prec
    x = torch.randn(n, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    ms = timeit(lambda: model.inverse(x))
    print(f"MAF x9 D=64 density {prec}: rows={n} {ms:.3f} ms  {n/ms/1e3:.2f} Mrows/s  ({516*n/ms/1e6:.0f} GB/s algorithmic)", flush=True)

# Sequential directions: sampling through MAF (D MADE passes per flow) and NSF_AR.forward
for f in model.flows:
    f.precision = "fp32"
for n in (1 << 12, 1 << 16):
    z = torch.randn(n, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    ms = timeit(lambda: model.forward(z), iters=3)
    print(f"MAF x9 D=64 sampling (64 sequential passes per flow): rows={n} {ms:.3f} ms  {n/ms:.1f} krows/s", flush=True)

g2 = load_golden("nsfar2_d3")
ar = load_flow_model(golden_spec(g2), golden_sd(g2), return_intermediates=False)
n = 1 << 20
z = torch.randn(n, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
for name, fn in (("forward (sequential)", lambda: ar.forward(z)), ("inverse", lambda: ar.inverse(z))):
    ms = timeit(fn)
    print(f"ActNorm+Glow+NSF_AR x2 D=3 {name}: rows={n} {ms:.3f} ms  {n/ms/1e3:.2f} Mrows/s", flush=True)
