"""MNFLinear(4096, 4096).kl_div() a few times (for ncu launch lists of the KL path, BASELINE config 5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)
from torch_mnf.layers import MNFLinear

torch.manual_seed(0)
layer = MNFLinear(4096, 4096).cuda()
for _ in range(int(os.environ.get("ITERS", 3))):
    kl = layer.kl_div()
torch.cuda.synchronize()
print(float(kl))
