import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
from torch_mnf.layers import MNFLinear
from torch_mnf.layers._mnf_ops import Noise
torch.manual_seed(0)
layer = MNFLinear(4096, 4096).cuda(); layer.precision = "tf32"
x = torch.randn(64, 4096, device="cuda")
S = int(os.environ.get("S", 1024))
for _ in range(2):
    out = layer.forward_mc(x, S, noise=Noise(None, x.device, 0, seed=1))
torch.cuda.synchronize(); print(out.shape, float(out.std()))
