#!/usr/bin/env python
"""bench.py -- flow log-prob throughput on the BASELINE.json headline configuration.

Workload ("cfg2"): [ActNormFlow, Glow, NSF_CL(K=8, B=3, n_h=16)] x 3 on synthetic 2-D points
x = 1.5 * randn(2^24, 2) (seed 0), one step = log p(x) for every point of the batch
(inverse pass + log-det + standard-normal base density, the quantity tests/test_flows.py:22-24
of the reference assembles).  One process per GPU; under torchrun each rank owns its own 2^24
points (weak scaling) and the per-point log-probs are all-gathered over NCCL, chunk by chunk,
overlapped with the next chunk's kernel.

    python bench.py --gpus 1 --steps 20 --warmup 3            # our arm
    python bench.py --impl reference --steps 3 --warmup 1      # CPU oracle port on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the meaning of every key.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "torch-mnf_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_POINTS = 1 << 24
K_BINS, BOUND, N_H = 8, 3, 16
BYTES_PER_POINT = 12  # 8 B point in + 4 B log-prob out (z is not materialised in log-prob mode)
MLP_FMA_PER_POINT = 3 * 2 * (16 + 256 + 256 + 16 * 23)  # 5376 fused multiply-adds in the conditioners
NCU_TRAFFIC_BYTES_PER_STEP = 1895.2e6  # dram read+write summed over the 6 segment launches of one step at 2^24 points (ncu --set full, r01)
METRIC = "flow log-prob points/s"
WORKLOAD = "cfg2: [ActNormFlow, Glow, NSF_CL(K=8,B=3,n_h=16)] x3, 2-D points, batch 2^24 per GPU"


def specs():
    return [
        {"type": "ActNormFlow", "dim": 2, "scale": True, "shift": True},
        {"type": "Glow", "dim": 2},
        {"type": "NSF_CL", "dim": 2, "K": K_BINS, "B": BOUND, "n_h": N_H},
    ] * 3


def make_points(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 1.5 * torch.randn(n, 2, generator=g)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json (measured)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.06] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        mx = max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(rows)}


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ---------------------------------------------------------------------------------------
def cpu_model_state():
    """Weights for the CPU arm: same construction as the CUDA arm (seed 0, ActNorm init on the
    first 4096 points), produced by the oracle alone so the reference arm needs no GPU."""
    from tests.helpers import random_flow_sd

    from oracle import flows_cpu

    sp = specs()
    sd = random_flow_sd(sp, seed=0, scale=0.6)
    x0 = make_points(4096)
    v = x0
    for i in reversed(range(len(sp))):  # data-dependent init in inverse order (core.py:30)
        p = flows_cpu.sub(sd, f"flows.{i}.")
        if sp[i]["type"] == "ActNormFlow":
            s, t = flows_cpu.actnorm_init(p, sp[i], v)
            sd[f"flows.{i}.s"], sd[f"flows.{i}.t"] = s, t
            p = flows_cpu.sub(sd, f"flows.{i}.")
        v, _ = flows_cpu.apply_flow(p, sp[i], v, True)
    return sp, sd


def cpu_time_sample(sp, sd, n, reps=1):
    from oracle import flows_cpu

    x = make_points(n, seed=1)
    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            flows_cpu.log_prob(sd, sp, x)
            best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sp, sd = cpu_model_state()
    probe = cpu_time_sample(sp, sd, 1 << 16)
    budget = 90.0 / max(1, args.steps + args.warmup)
    n = 1 << 16
    while n < (1 << 22) and probe * (2 * n / (1 << 16)) < budget:
        n *= 2
    for _ in range(args.warmup):
        cpu_time_sample(sp, sd, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_time_sample(sp, sd, n)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    sample = f"{n} points per step (2^{n.bit_length() - 1}) of the 2^24-point workload, torch CPU, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def build_model(device):
    from tests.helpers import load_flow_model, random_flow_sd

    sp = specs()
    model = load_flow_model(sp, random_flow_sd(sp, seed=0, scale=0.6), device=device, return_intermediates=False)
    for f in model.flows:  # re-arm the data-dependent init, as a fresh model would have it
        if hasattr(f, "data_dep_init_done"):
            f.data_dep_init_done = False
    model.inverse(make_points(4096).to(device))
    return model


def pin_to_gpu_numa_node(local: int) -> str:
    """Bind this rank's host threads to the cores next to its GPU before any pinned buffer is allocated (first touch
    then places the buffers on that NUMA node): with 8 ranks the host-to-host figure is limited by cross-socket
    traffic otherwise.  Best effort; returns a note for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores local to GPU {local}"
    except Exception as e:  # noqa: BLE001
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


def run_ours(args):
    import torch.distributed as dist
    from torch_mnf import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_numa_node(local) if world > 1 else "unchanged (single rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(dev)
    n = args.points
    chunks = args.chunks if world > 1 else 1
    assert n % (2 * chunks) == 0
    cn = n // chunks
    x_host = make_points(n, seed=100 + rank).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    # Result gather.  Preferred: NVLink peer memory -- the kernel stores each log-prob into every rank's gather
    # buffer while it computes (torch_mnf.distributed.PeerGather: NVLS multicast or per-peer stores), so the step
    # has no data-moving collective, only barriers.  Fallback (--gather nccl, or no symmetric memory): chunked
    # NCCL all-gather on a side stream; the kernel writes this rank's slice of the gather buffer directly.
    peer = None
    if world > 1 and args.gather == "peer":
        try:
            from torch_mnf.distributed import PeerGather

            peer = PeerGather(n, dev)
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] peer-memory gather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
            peer = None
    if peer is not None:
        chunks = 1
        cn = n
        gathered = peer.buffer.view(1, world, n)
        gather_out = peer.gather_out(use_multicast=not args.no_multicast)
    else:
        gathered = torch.empty((chunks, world, cn), device=dev, dtype=torch.float32)
        gather_out = None
    comm = torch.cuda.Stream(device=dev) if (world > 1 and peer is None) else None

    def step():
        cur = torch.cuda.current_stream(dev)
        if peer is not None:
            peer.barrier()  # peers are done reading the previous step's results
            model.log_prob(x, out=gathered[0, rank], gather=gather_out)
            peer.barrier()  # every rank's stores have landed everywhere
            return
        for c in range(chunks):
            model.log_prob(x[c * cn:(c + 1) * cn], out=gathered[c, rank])
            if world > 1:
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(gathered[c].view(-1), gathered[c, rank])
        if world > 1:
            cur.wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    launches0 = _lib.lib().mnf_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.lib().mnf_launch_count() - launches0  # kernels launched by libmnf_b200.so in the timed region
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        tmax = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    gather_ok = True
    if world > 1:
        # every rank must now hold every rank's log-probs: compare a checksum per slice with an all-gather of the
        # owners' checksums
        mine = gathered.view(-1, world, cn)[:, rank].double().sum().reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        seen = torch.stack([gathered.view(-1, world, cn)[:, r].double().sum() for r in range(world)])
        gather_ok = bool(torch.allclose(seen, torch.cat(sums), rtol=1e-9, atol=1e-6))
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- kernel-only duration (same launches, no collective) for the roofline ----
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        model.log_prob(x, out=gathered.view(-1)[:n])
        b.record()
    torch.cuda.synchronize(dev)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev)

    # ---- end to end through the module API with HOST buffers (pinned), copies inside ----
    out_host = torch.empty(n, dtype=torch.float32).pin_memory()
    e_chunks = args.e2e_chunks
    ecn = n // e_chunks
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    xd = [torch.empty((ecn, 2), device=dev) for _ in range(2)]
    od = [torch.empty(ecn, device=dev) for _ in range(2)]

    # chunk buffers are double-buffered ACROSS steps too: the first copy-in of step k+1 overlaps the last kernels and
    # copy-outs of step k (every step still moves all of its input H2D and all of its output D2H inside the timed region)
    ready = [None, None]  # copy-out of the chunk that last used od[b] has finished
    freed = [None, None]  # kernel that last read xd[b] has finished

    def e2e_step():
        cur = torch.cuda.current_stream(dev)
        for c in range(e_chunks):
            b = c & 1
            with torch.cuda.stream(s_in):
                if freed[b] is not None:
                    s_in.wait_event(freed[b])
                xd[b].copy_(x_host[c * ecn:(c + 1) * ecn], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(s_in)
            cur.wait_event(e_in)
            if ready[b] is not None:
                cur.wait_event(ready[b])
            model.log_prob(xd[b], out=od[b])
            e_k = torch.cuda.Event()
            e_k.record(cur)
            freed[b] = e_k
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_k)
                out_host[c * ecn:(c + 1) * ecn].copy_(od[b], non_blocking=True)
                e_o = torch.cuda.Event()
                e_o.record(s_out)
            ready[b] = e_o

    def e2e_join():
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)

    for _ in range(2):
        e2e_step()
    e2e_join()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_join()
    b.record()
    barrier()
    e_ms = a.elapsed_time(b) / args.steps
    if world > 1:
        tmax = torch.tensor([e_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e_ms = float(tmax)
    e2e_value = world * n / (e_ms * 1e-3)
    # sanity: the e2e path produced finite log-probs equal to the resident path
    chk = torch.allclose(out_host[:4096], gathered.view(-1)[:4096].cpu(), rtol=1e-6, atol=1e-6) if chunks == 1 else True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, _ = peaks()
    achieved = BYTES_PER_POINT * n / (k_ms * 1e-3) / 1e9
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sms * 128 * 2 * sm_clock * 1e6 / 1e12
    mlp_tflops = 2 * MLP_FMA_PER_POINT * n / (k_ms * 1e-3) / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": n, "l2": "inputs (134 MB/GPU) larger than the 126 MB L2",
                   "collective": ("none (N=1)" if world == 1 else
                                  ("fused in the kernel: log-probs stored to every rank over NVLink peer memory ("
                                   + ("NVLS multicast multimem.st" if gather_out.multicast_ptr else "per-peer st.global")
                                   + "), 2 barriers per step") if peer is not None else
                                  f"NCCL all_gather of log-probs in {chunks} chunks overlapped with compute"), "parallelism": f"points sharded over {world} GPU(s), weights replicated"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * n,
                "ms_per_step": e_ms, "matches_resident_path": bool(chk), "host_affinity": affinity,
                "how": f"pinned host -> {e_chunks} chunks double-buffered over 3 streams (pipelined across steps) -> NormalizingFlowModel.log_prob -> pinned host"},
        "gpu_launches": launches, "gather_verified": gather_ok,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC_BYTES_PER_STEP * n / N_POINTS,
                     "kernel": "flow_cbank_kernel<16,8> x6 segments (+ cbank_stage_kernel)",
                     "kernel_ms": k_ms, "bytes_per_point": BYTES_PER_POINT, "peak_source": peak_src,
                     "note": "per STEP (6 segment launches): algorithmic 12 B/pt; the segmented stack moves ~113 B/pt "
                             "(ncu, profiles/r01_flow_cbank_ncu_full.md); kernel is fp32-FMA-pipe bound, not HBM "
                             "bound (DESIGN.md): see fma_pipe"},
        "fma_pipe": {"mlp_tflops": mlp_tflops, "fp32_peak_tflops_at_sampled_clock": fp32_peak,
                     "frac_mlp_only": mlp_tflops / fp32_peak, "fma_per_point_mlp": MLP_FMA_PER_POINT,
                     "note": "conditioner-MLP FMAs only; spline arithmetic shares the same pipe"},
    }
    if world == 1:
        # The conditioner-free part of the same stack ([ActNormFlow, Glow] x 3) is the flow workload that IS HBM-bound:
        # reported next to the headline so that the HBM roofline of the streaming path can be read off this line too.
        try:
            from tests.helpers import load_flow_model, random_flow_sd

            sub_specs = [s_ for s_ in specs() if s_["type"] != "NSF_CL"]
            sub = load_flow_model(sub_specs, random_flow_sd(sub_specs, seed=0), device=dev, return_intermediates=False)
            ys, lds = torch.empty_like(x), torch.empty(n, device=dev)
            prog = sub._program()
            for _ in range(3):
                prog.run(x, True, out=ys, log_det=lds)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(20):
                prog.run(x, True, out=ys, log_det=lds)
            b_.record()
            torch.cuda.synchronize(dev)
            sub_ms = a_.elapsed_time(b_) / 20
            out["hbm_bound_substack"] = {
                "workload": "[ActNormFlow, Glow] x3 (the conditioner-free flows of the headline stack), inverse with z and log_det stored",
                "bytes_per_point": 20, "ms": sub_ms, "achieved": 20 * n / (sub_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": 20 * n / (sub_ms * 1e-3) / 1e9 / hbm_peak, "kernel": "affine_stream_kernel"}
            del ys, lds, sub
        except Exception as e:  # noqa: BLE001  (diagnostic extra, never fails the bench)
            out["hbm_bound_substack"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sp, sd = cpu_model_state()
        probe = cpu_time_sample(sp, sd, 1 << 16)
        ncpu = 1 << 16
        while ncpu < (1 << 22) and probe * (2 * ncpu / (1 << 16)) < 8.0:
            ncpu *= 2
        t = cpu_time_sample(sp, sd, ncpu, reps=2)
        out["cpu_baseline"] = {"value": ncpu / t, "unit": "points/s", "cores": cores, "kind": "port",
                               "sample": f"{ncpu} points (2^{ncpu.bit_length() - 1}), best of 2, torch CPU oracle port, {cores} threads"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
# secondary workload (not the default line): BASELINE config 4, MNF-LeNet MC prediction
# ---------------------------------------------------------------------------------------
def run_lenet(args):
    """MNF-LeNet Monte-Carlo predictive samples/s (BASELINE config 4).  1024 synthetic 28x28 images; every rank
    draws `--mc-samples` samples per image per step (weak scaling: the MC-sample axis is sharded, conv z shared
    through the common seed, per-row noise keyed by the global row), reduces them to per-image class
    probabilities locally and all-reduces the [1024, 10] sums -- the raw samples never cross NVLink."""
    import torch.distributed as dist
    from tests.helpers import golden_sd, load_golden
    from torch_mnf import _lib
    from torch_mnf.distributed import reduce_mc_probs
    from torch_mnf.models import MNFLeNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = MNFLeNet()
    net.load_state_dict(golden_sd(load_golden("mnf_lenet")))  # the briefly trained fixture (tests/golden)
    net.to(dev)
    n_img, S, chunk = 1024, args.mc_samples, 32
    imgs_host = torch.rand(n_img, 1, 28, 28, generator=torch.Generator().manual_seed(0)).pin_memory()
    imgs = imgs_host.to(dev)
    probs_host = torch.empty(n_img, 10).pin_memory()

    def step(step_idx, from_host=False):
        x = imgs_host.to(dev, non_blocking=True) if from_host else imgs
        sums = torch.zeros(n_img, 10, device=dev)
        for c in range(0, S, chunk):
            s_here = min(chunk, S - c)
            lp = net(x, n_samples=s_here, seed=1000 + step_idx * 131 + c, row_offset=(rank * S + c) * n_img)
            sums += lp.exp().view(s_here, n_img, 10).sum(0)
        if world > 1:
            dist.all_reduce(sums)
        probs = sums / (S * world)
        if from_host:
            probs_host.copy_(probs, non_blocking=True)
        return probs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    l0 = _lib.lib().mnf_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    a.record()
    for i in range(args.steps):
        probs = step(100 + i)
    b.record()
    barrier()
    t1 = time.time()
    launches = _lib.lib().mnf_launch_count() - l0
    ms = a.elapsed_time(b) / args.steps
    a.record()
    for i in range(args.steps):
        step(200 + i, from_host=True)
    b.record()
    barrier()
    e_ms = a.elapsed_time(b) / args.steps
    if world > 1:
        tt = torch.tensor([ms, e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e_ms = float(tt[0]), float(tt[1])
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        total = world * n_img * S
        flops = 8.2e6 * total  # ~8.2 MFLOP per sample (dense count, conv1 per sample)
        out = {
            "metric": "MNF-LeNet MC predictive samples/s", "value": total / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 tensor cores (conv2, fc1) / f32",
            "data": "synthetic",
            "config": {"workload": f"cfg4: MNF-LeNet (696,950 params), 1024 images x {S} MC samples per GPU per step",
                       "parallelism": f"MC samples sharded over {world} GPU(s); all_reduce of [1024,10] probability sums",
                       "l2": "activations (46 KB/sample before the pool of conv1, 11.5 KB/sample after it) far exceed L2"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": total / (e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": n_img * 784 * 4,
                    "d2h_bytes_per_step": n_img * 40, "ms_per_step": e_ms},
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12 / world, "peak": peaks()[2].get("bf16_tflops_sustained", 1400.0) / 2,
                         "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / world / (peaks()[2].get("bf16_tflops_sustained", 1400.0) / 2),
                         "traffic": None, "note": "dense-equivalent 8.2 MFLOP/sample against the TF32 peak (= half the measured bf16 "
                                                  "sustained peak); the pipeline is bound by operand generation in shared memory (implicit-GEMM conv2) and Philox noise, not by the tensor pipe"},
            "cpu_baseline": None,
        }
        if world == 1 and not args.no_cpu:
            # the reference's CPU path (oracle port: same ATen ops) on a bounded sample: 8 images x 250 MC samples
            from oracle import mnf_cpu
            from oracle.noise import FreshNoise

            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            xc = imgs_host[:8].repeat(250, 1, 1, 1)
            best = float("inf")
            with torch.no_grad():
                for _ in range(3):
                    t_0 = time.perf_counter()
                    mnf_cpu.lenet_forward(sd_cpu, xc, FreshNoise())
                    best = min(best, time.perf_counter() - t_0)
            out["cpu_baseline"] = {"value": xc.size(0) / best, "unit": "samples/s", "cores": cores, "kind": "port",
                                   "sample": "8 images x 250 MC samples (2000 rows), best of 3"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS, help="points per GPU (default 2^24)")
    ap.add_argument("--chunks", type=int, default=4, help="all-gather chunks per step when N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-chunks", type=int, default=2, help="chunks per step of the host-to-host (e2e) pipeline")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1 result gather: fused peer-memory stores (default) or NCCL all-gather")
    ap.add_argument("--no-multicast", action="store_true", help="peer gather: per-peer stores even if NVLS multicast exists")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 (default, the headline flow log-prob line) or cfg4 (MNF-LeNet MC prediction)")
    ap.add_argument("--mc-samples", type=int, default=64, help="cfg4: MC samples per image per GPU per step")
    args = ap.parse_args()
    torch.set_grad_enabled(False)  # the workload is density evaluation / prediction, not training
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_lenet(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
