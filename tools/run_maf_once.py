"""Runs the cfg3 MAF x9 dim-64 density pass a few times (for ncu captures of made_fast_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden

g = load_golden("maf9_d64")
model = load_flow_model(golden_spec(g), golden_sd(g), return_intermediates=False)
for f in model.flows:
    f.precision = os.environ.get("PREC", "fp32")
n = int(os.environ.get("N", 1 << 20))
x = torch.randn(n, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
sampling = os.environ.get("DIR", "inverse") == "forward"  # forward = the 64-pass sampling direction
for _ in range(int(os.environ.get("ITERS", 2))):
    zs, ld = model.forward(x) if sampling else model.inverse(x)
torch.cuda.synchronize()
print("done", float(ld.mean()))
