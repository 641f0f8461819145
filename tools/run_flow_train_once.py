"""A few training steps of ActNorm + NSF_CL x2 on 128 moons points (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch
from torch.distributions import MultivariateNormal

import torch_mnf.flows as nf
from torch_mnf import data

n = int(os.environ.get("N", 128))
torch.manual_seed(0)
flows = [nf.ActNormFlow(dim=2), nf.NSF_CL(dim=2, K=8, B=3, n_h=16), nf.ActNormFlow(dim=2), nf.NSF_CL(dim=2, K=8, B=3, n_h=16)]
model = nf.NormalizingFlowModel(MultivariateNormal(torch.zeros(2), torch.eye(2)), flows).cuda()
x = data.sample_moons(n).cuda()
adam = torch.optim.Adam(model.parameters())
import time
for i in range(int(os.environ.get("ITERS", 6))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, ld = model.inverse(x)
    loss = -(ld + model.base_log_prob(x)).sum()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    model.zero_grad()
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    adam.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"step {i}: forward {1e3*(t1-t0):.2f} ms, backward {1e3*(t2-t1):.2f} ms, adam {1e3*(t3-t2):.2f} ms  loss {float(loss.detach()):.3f}", flush=True)
