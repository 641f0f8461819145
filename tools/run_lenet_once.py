import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
from tests.helpers import golden_sd, load_golden
from torch_mnf.models import MNFLeNet
g = load_golden("mnf_lenet")
net = MNFLeNet(); net.load_state_dict(golden_sd(g)); net.cuda()
imgs = torch.rand(1024, 1, 28, 28, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3):
    y = net(imgs, n_samples=25)
torch.cuda.synchronize(); print(y.shape)
