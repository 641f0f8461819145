"""Small run of the table kernel (csrc/flow_pl.cu) for compute-sanitizer: builder + point kernel on the two BASELINE dim-2
stacks, a wide stack whose tables exceed shared memory and one that overflows its table (fp32 fallback)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)
from tests.helpers import load_flow_model, random_flow_sd
from tests.test_flow_pl_gpu import SHAPES
from tests.test_flows_gpu import ORACLE_CASES

cases = {"cfg2": ORACLE_CASES["cfg2_shape"], "cfg1": ORACLE_CASES["cfg1_shape"], "nsf_h64_x3": SHAPES["nsf_h64_x3"],
         "affine_overflow": SHAPES["affine_overflow"]}
for name, specs in cases.items():
    model = load_flow_model(specs, random_flow_sd(specs, seed=0, scale=0.5))
    x = torch.randn(int(os.environ.get("N", 3000)), 2, device="cuda")
    for inverse in (True, False):
        y, ld, _, lp = model._program().run(x, inverse, want_base_lp=True, kernel=6)
    lp2 = model.log_prob(x)  # staged tables
    torch.cuda.synchronize()
    print(name, float(lp2.nan_to_num().mean()))
