"""Timing + error statistics of the tensor-core spline kernel (variant 4) next to the shared-memory FFMA2 kernel
(variant 2; round 1's constant-bank variant 3 was removed in round 2) on BASELINE config 2.  Run on a GPU box:  python tests/measure/flow_tc_timing.py [log2_rows]
Env: MNF_FTC_V (groups * 10 + tiles per thread, + 100 = biases in the epilogue), MNF_FTC_DEBUG (1 no spline, 2 no MMAs, 4 truncating hi split)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

from tests.helpers import load_flow_model, random_flow_sd

torch.set_grad_enabled(False)
specs = [{"type": "ActNormFlow", "dim": 2, "scale": True, "shift": True}, {"type": "Glow", "dim": 2},
         {"type": "NSF_CL", "dim": 2, "K": 8, "B": 3, "n_h": 16}] * 3
sd = random_flow_sd(specs, seed=0, scale=0.6)
model = load_flow_model(specs, sd, device="cuda:0", return_intermediates=False)
prog = model._program()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
x = 1.5 * torch.randn(n, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
out = {"rows": n, "V": os.environ.get("MNF_FTC_V"), "debug": os.environ.get("MNF_FTC_DEBUG")}
lp_out = torch.empty(n, device="cuda")
for kernel in (2, 4):
    for _ in range(3):
        prog.run(x, True, log_prob_only=True, kernel=kernel, log_prob_out=lp_out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        prog.run(x, True, log_prob_only=True, kernel=kernel, log_prob_out=lp_out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"k{kernel}_ms"] = ms
    out[f"k{kernel}_gpts"] = n / ms / 1e6
    out[f"k{kernel}_lp"] = lp_out[: 1 << 16].clone()
d = (out.pop("k2_lp") - out.pop("k4_lp")).abs()
out["max_abs_diff_k2_k4"] = float(d.max())
out["mean_abs_diff_k2_k4"] = float(d.mean())
print(json.dumps(out))
