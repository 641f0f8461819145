"""Throughput of the tcgen05 TF32 GEMM and the MNFLinear tensor-core forward (config-5 shapes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
import torch_mnf.layers
from torch_mnf import _lib
from torch_mnf.layers import MNFLinear
from torch_mnf.layers._mnf_ops import Noise


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)


for M, N, K in [(8192, 4096, 4096), (65536, 4096, 4096)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 64; out = torch.empty(M, N, device="cuda")
    f = lambda: _lib.check(_lib.lib().mnf_tc_linear(A.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, 0, 0, _lib.stream_ptr(A.device)), "tc")
    ms = timeit(f)
    print(f"tc_linear M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = timeit(lambda: torch.matmul(A, W.T, out=out))
    print(f"  cuBLAS tf32 (library reference): {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    del A, W, out

torch.manual_seed(0)
layer = MNFLinear(4096, 4096).cuda()
x = torch.randn(64, 4096, device="cuda")
for S in (128, 1024):
    R = 64 * S
    for prec in ("tf32",):
        layer.precision = prec
        ms = timeit(lambda: layer.forward_mc(x, S, noise=Noise(None, x.device, 0, seed=1)), iters=3)
        print(f"MNFLinear.forward_mc rows={R} {prec}: {ms:.2f} ms  {R/ms*1e3/1e6:.3f} Mrows/s  main-GEMM {2*R*4096*4096/ms/1e9:.1f} TFLOP/s", flush=True)
