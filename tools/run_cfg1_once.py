import os, sys
ROOT = "/root/repo"
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden
torch.set_grad_enabled(False)
g = load_golden("rnvp9_moons")
model = load_flow_model(golden_spec(g), golden_sd(g), device="cuda:0", return_intermediates=False)
x = torch.randn(4096, 2, device="cuda")
f = model.log_prob_fn(4096)
for _ in range(5):
    f(x)
torch.cuda.synchronize()
