"""Config-3 density (MAF x9, D=64, 2^20 rows): fused tcgen05 kernel (2 / 3 tiles in flight, z + log_det and log-prob-only
modes) next to the exact-fp32 kernel.  CUDA events, inputs (268 MB) larger than L2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden
from torch_mnf.layers import made

g = load_golden("maf9_d64")
model = load_flow_model(golden_spec(g), golden_sd(g), return_intermediates=False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
x = torch.randn(n, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))


def t_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {}
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [41, 32, 31, 21]
for tiles in variants:
    made.VARIANT = tiles
    ms = t_ms(lambda: model.inverse(x))
    ms_lp = t_ms(lambda: model.log_prob(x))
    out[f"fused_v{tiles}"] = {"ms": ms, "rows_per_s": n / ms * 1e3, "GB_s_516B": 516 * n / ms / 1e6,
                              "dense_tflops": 103680 * n / ms / 1e9, "log_prob_ms": ms_lp}
made.VARIANT = 0
for f in model.flows:
    f.precision = "fp32"
ms = t_ms(lambda: model.inverse(x))
out["exact_fp32_made_fast"] = {"ms": ms, "rows_per_s": n / ms * 1e3}
print(json.dumps(out, indent=1))
