"""Latency of NormalizingFlowModel.log_prob on BASELINE config 1 (RNVP x9) at small batch sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[os.path.join(ROOT,"torch-mnf_b200"), ROOT]
import torch
torch.set_grad_enabled(False)
from tests.helpers import golden_sd, load_flow_model, load_golden, t
from tests.test_flows_gpu import ORACLE_CASES
specs=ORACLE_CASES["cfg1_shape"]; g=load_golden("rnvp9_moons")
model=load_flow_model(specs, golden_sd(g), return_intermediates=False)
for n in (256, 4096, 16384):
    x=t(g,"inv/x").repeat(200,1)[:n].contiguous().cuda()
    for _ in range(5): model.log_prob(x)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): model.log_prob(x)
    b.record(); torch.cuda.synchronize()
    print(f"batch {n}: {a.elapsed_time(b) / 200 * 1e3:.1f} us/call", flush=True)

# floor of a call: the same API on a stack with no conditioner nets (nothing to stage, a few FMAs per point)
import torch_mnf.flows as nf
from torch.distributions import MultivariateNormal

tiny = nf.NormalizingFlowModel(MultivariateNormal(torch.zeros(2), torch.eye(2)), [nf.AffineConstantFlow(2)]).cuda()
x = torch.randn(4096, 2, device="cuda")
for _ in range(5):
    tiny.log_prob(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    tiny.log_prob(x)
b.record(); torch.cuda.synchronize()
print(f"call floor (one AffineConstantFlow, batch 4096): {a.elapsed_time(b) / 200 * 1e3:.1f} us/call", flush=True)
