"""kl_div() call time: MNFLinear(4096, 4096) (BASELINE config 5) and MNFLeNet (four layers).  Run on a GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

from torch_mnf import _lib
from torch_mnf.layers import MNFLinear
from torch_mnf.models import MNFLeNet

torch.set_grad_enabled(False)
torch.manual_seed(0)
out = {}
for name, mod in (("MNFLinear(4096,4096).kl_div", MNFLinear(4096, 4096).cuda()), ("MNFLeNet.kl_div", MNFLeNet().cuda())):
    for _ in range(5):
        mod.kl_div()
    torch.cuda.synchronize()
    _lib.launch_stats(reset=True)
    mod.kl_div()
    sites = _lib.launch_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        mod.kl_div()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"ms_per_call": e0.elapsed_time(e1) / 50, "launches": sites}
ms = out["MNFLinear(4096,4096).kl_div"]["ms_per_call"]
out["MNFLinear(4096,4096).kl_div"]["weights_GBps_whole_call"] = 2 * 4096 * 4096 * 4 / (ms * 1e-3) / 1e9
print(json.dumps(out))
