"""cProfile of the host side of a config-1 module call (NormalizingFlowModel.log_prob, 4096 points)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden
torch.set_grad_enabled(False)
g = load_golden("rnvp9_moons")
model = load_flow_model(golden_spec(g), golden_sd(g), device="cuda:0", return_intermediates=False)
x = torch.randn(4096, 2, device="cuda")
lp = torch.empty(4096, device="cuda")
for _ in range(50):
    model.log_prob(x, out=lp)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    model.log_prob(x, out=lp)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
