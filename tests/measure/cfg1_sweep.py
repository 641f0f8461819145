"""Config-1 stack (RNVP x9, h = 24): log_prob call time by batch size for the lane-split kernel (variant 5), the
one/two-point-per-thread kernel (variant 2) and the bound-handle call.  Run on a GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden

torch.set_grad_enabled(False)
g = load_golden("rnvp9_moons")
model = load_flow_model(golden_spec(g), golden_sd(g), device="cuda:0", return_intermediates=False)
prog = model._program()


def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


out = {"MNF_LANES": os.environ.get("MNF_LANES")}
for rows in (256, 1024, 4096, 8192, 16384, 32768):
    x = torch.randn(rows, 2, device="cuda")
    lp = torch.empty(rows, device="cuda")
    r = {}
    for k in (2, 5):
        r[f"variant{k}_us"] = timed(lambda: prog.run(x, True, log_prob_only=True, kernel=k, log_prob_out=lp))
    f = model.log_prob_fn(max_rows=rows)
    r["bound_us"] = timed(lambda: f(x, lp))
    # kernel-only: many calls back to back are GPU-bound when the kernel is longer than the host path
    out[rows] = r
print(json.dumps(out))
