"""Copy-only ceiling of bench.py's end-to-end figure: per step and GPU, 2^24 points x 8 B host -> device and 2^24 x 4 B
device -> host from / to pinned memory, both directions concurrently on two streams, NO kernel.  Run alone or under
torchrun (one rank per GPU); rank 0 prints one JSON line with the aggregate rate in the e2e metric's unit."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 24
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
x_host = torch.empty(n, 2).pin_memory()
o_host = torch.empty(n).pin_memory()
xd, od = torch.empty(n, 2, device=dev), torch.zeros(n, device=dev)
s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def run(k, h2d=True, d2h=True):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s_in.wait_stream(torch.cuda.current_stream(dev))
    s_out.wait_stream(torch.cuda.current_stream(dev))
    for _ in range(k):
        if h2d:
            with torch.cuda.stream(s_in):
                xd.copy_(x_host, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                o_host.copy_(od, non_blocking=True)
    torch.cuda.current_stream(dev).wait_stream(s_in)
    torch.cuda.current_stream(dev).wait_stream(s_out)
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / k
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


run(3)
both, only_in, only_out = run(steps), run(steps, d2h=False), run(steps, h2d=False)
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "points_per_gpu_per_step": n, "ms_per_step_both_directions": both,
        "copy_only_ceiling_points_per_s": world * n / (both * 1e-3),
        "h2d_GBps_per_gpu_concurrent": 8 * n / (both * 1e-3) / 1e9, "d2h_GBps_per_gpu_concurrent": 4 * n / (both * 1e-3) / 1e9,
        "h2d_alone_ms": only_in, "h2d_alone_GBps_per_gpu": 8 * n / (only_in * 1e-3) / 1e9,
        "d2h_alone_ms": only_out, "d2h_alone_GBps_per_gpu": 4 * n / (only_out * 1e-3) / 1e9,
        "aggregate_host_GBps_both": world * 12 * n / (both * 1e-3) / 1e9}))
if world > 1:
    dist.destroy_process_group()
