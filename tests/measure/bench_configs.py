"""All five BASELINE.json configurations: CUDA path (this repo) next to the CPU oracle port on the box's
host cores.  Prints one JSON object per config and writes gpurun_out/configs.json + configs.md."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-mnf_b200"), ROOT]
import torch

torch.set_grad_enabled(False)  # inference kernels; training goes through mnf_flow_stack_backward
from oracle import flows_cpu, mnf_cpu
from oracle.noise import FreshNoise
from tests.helpers import golden_sd, golden_spec, load_flow_model, load_golden, random_flow_sd, t
from tests.test_flows_gpu import ORACLE_CASES

dev = torch.device("cuda")
cores = os.cpu_count()
torch.set_num_threads(cores)
results = []


def gpu_time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[0], ts[len(ts) // 2]


def cpu_time(fn, reps=2):
    best = 1e30
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best


def record(**kw):
    results.append(kw)
    print(json.dumps(kw), flush=True)


# ---- cfg1: RNVP x9 on half-moons-like 2-D points, batch 4096 ----
specs = ORACLE_CASES["cfg1_shape"]
g = load_golden("rnvp9_moons")
sd = golden_sd(g)
model = load_flow_model(specs, sd, return_intermediates=False)
x = t(g, "inv/x").repeat(16, 1)[:4096].contiguous()
xd = x.cuda()
best, med = gpu_time(lambda: model.log_prob(xd), iters=50)
cpu = cpu_time(lambda: flows_cpu.log_prob(sd, specs, x), reps=5)
record(config="cfg1 RNVP x9 (AffineHalfFlow), batch 4096, log_prob", gpu_us_per_call=med * 1e3, gpu_points_per_s=4096 / (med * 1e-3),
       cpu_points_per_s=4096 / cpu, cpu_cores=cores, note="latency-bound inside the kernel: 2048 threads on 16 CTAs (a CUDA-graph replay takes the same time)")
from torch_mnf.graphs import graphed_inference
call = graphed_inference(lambda v: model.log_prob(v), (xd,))
best, med = gpu_time(lambda: call(xd), iters=200)
record(config="cfg1 same, replayed from a CUDA graph (torch_mnf.graphs.graphed_inference)", gpu_us_per_call=med * 1e3,
       gpu_points_per_s=4096 / (med * 1e-3), note="input copied into the graph's static buffer, one replay per call")
xb = x.repeat(4096, 1).cuda()[: 1 << 24].contiguous()
best, med = gpu_time(lambda: model.log_prob(xb), iters=5)
record(config="cfg1 stack at batch 2^24 (throughput regime)", gpu_ms=best, gpu_points_per_s=(1 << 24) / (best * 1e-3),
       fp32_tflops_mlp=2 * 21600 * (1 << 24) / (best * 1e-3) / 1e12)
del xb

# ---- cfg3: MAF x9, D = 64, density ----
g = load_golden("maf9_d64")
sd3, specs3 = golden_sd(g), golden_spec(g)
model3 = load_flow_model(specs3, sd3, return_intermediates=False)
n = 1 << 20
x3 = torch.randn(n, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
best, med = gpu_time(lambda: model3.inverse(x3), iters=5)
xc = torch.randn(1 << 16, 64)
cpu = cpu_time(lambda: flows_cpu.stack(sd3, specs3, xc, True))
record(config="cfg3 MAF x9 D=64 density, batch 2^20 (exact-fp32 shared-memory MADE kernel, the default)", gpu_ms=best, gpu_rows_per_s=n / (best * 1e-3),
       hbm_gbs_algorithmic=516 * n / (best * 1e-3) / 1e9, cpu_rows_per_s=(1 << 16) / cpu, cpu_sample="2^16 rows", cpu_cores=cores)
for f in model3.flows:
    f.precision = "tf32"
best, med = gpu_time(lambda: model3.inverse(x3), iters=5)
record(config="cfg3 same, TF32 tensor-core GEMM chain (precision='tf32')", gpu_ms=best, gpu_rows_per_s=n / (best * 1e-3))
x3s = x3[: 1 << 16].contiguous()
best, med = gpu_time(lambda: model3._program().run(x3s, True, kernel="generic"), iters=3)
record(config="cfg3 same, generic interpreter (parity path)", gpu_ms=best, gpu_rows_per_s=(1 << 16) / (best * 1e-3), sample="2^16 rows")
for f in model3.flows:
    f.precision = "auto"
z3 = torch.randn(1 << 18, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
best, med = gpu_time(lambda: model3.forward(z3), iters=3)
record(config="cfg3 stack, sampling direction (MAF.forward: 64 sequential passes per flow), batch 2^18", gpu_ms=best,
       gpu_rows_per_s=(1 << 18) / (best * 1e-3), note="reference on CPU: 8.4 k rows/s (SURVEY.md 8f-2 probe)")
del x3, x3s, z3

# ---- cfg4: MNF-LeNet, 1024 images x 500 MC samples ----
from torch_mnf.models import MNFLeNet
g = load_golden("mnf_lenet")
sd4 = golden_sd(g)
net = MNFLeNet(); net.load_state_dict(sd4); net.cuda()
imgs = torch.rand(1024, 1, 28, 28, generator=torch.Generator().manual_seed(0))
imgs_d = imgs.cuda()
S_call = 25
fn = lambda: net(imgs_d, n_samples=S_call)
best, med = gpu_time(fn, iters=3, warm=1)
total_ms = best * (500 / S_call)
cpu = cpu_time(lambda: mnf_cpu.lenet_forward(sd4, imgs[:8].repeat(250, 1, 1, 1), FreshNoise()), reps=1)
record(config="cfg4 MNF-LeNet MC prediction, 1024 images x 500 samples", gpu_ms_total=total_ms, gpu_samples_per_s=1024 * 500 / (total_ms * 1e-3),
       per_call=f"{S_call} samples x 1024 images per call ({best:.1f} ms), {500 // S_call} calls",
       cpu_samples_per_s=2000 / cpu, cpu_sample="2000 rows", cpu_cores=cores, flops_per_sample=8.2e6,
       gpu_tflops=8.2e6 * 1024 * 500 / (total_ms * 1e-3) / 1e12)
best, med = gpu_time(lambda: net.kl_div(), iters=5)
record(config="cfg4 MNFLeNet.kl_div()", gpu_ms=med)
del net

# ---- cfg5: wide MNFLinear 4096 x 4096, 64 rows x 8192 MC samples + kl_div ----
from torch_mnf.layers import MNFLinear
from torch_mnf.layers._mnf_ops import Noise
torch.manual_seed(0)
layer = MNFLinear(4096, 4096).cuda()
x64 = torch.randn(64, 4096, device=dev)
S = int(os.environ.get("CFG5_SAMPLES", 8192))
R = 64 * S
fn = lambda: layer.forward_mc(x64, S, noise=Noise(None, dev, 0, seed=1))
best, med = gpu_time(fn, iters=2, warm=1)
sd5 = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
xc = x64.cpu().repeat(32, 1)
cpu = cpu_time(lambda: mnf_cpu.linear_forward(sd5, xc, FreshNoise()), reps=1)
record(config=f"cfg5 MNFLinear(4096,4096) forward, 64 rows x {S} MC samples (TF32 tensor cores)", gpu_ms=best,
       gpu_rows_per_s=R / (best * 1e-3), executed_tflops=(2 * R * 4096 * 4096 + 2 * 2 * R * 4096 * 64 * 3) / (best * 1e-3) / 1e12,
       main_gemm_tflops_equiv=2 * R * 4096 * 4096 / (best * 1e-3) / 1e12,
       cpu_rows_per_s=2048 / cpu, cpu_sample="2048 rows", cpu_cores=cores,
       note="variance GEMM evaluated once per distinct input row (64), not per sample")
best, med = gpu_time(lambda: layer.kl_div(), iters=5)
cpu = cpu_time(lambda: mnf_cpu.linear_kl_div(sd5, FreshNoise()), reps=2)
record(config="cfg5 MNFLinear(4096,4096).kl_div()", gpu_ms=med, cpu_ms=cpu * 1e3,
       hbm_gbs=2 * 4096 * 4096 * 4 / (med * 1e-3) / 1e9, note="reads W_mean + W_log_var once (134 MB), eps_w from Philox")

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
