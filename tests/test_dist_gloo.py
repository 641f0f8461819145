"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding + gather, and the MC
probability reduction.  The per-rank compute is stood in for by the oracle (this is a test)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "torch-mnf_b200"), root]
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import flows_cpu
    from tests.helpers import golden_sd, golden_spec, load_golden
    from torch_mnf.distributed import gather_rows, reduce_mc_probs, shard_range

    g = load_golden("nsfcl3_stack")
    sd, specs = golden_sd(g), golden_spec(g)
    gen = torch.Generator().manual_seed(0)
    x = 1.5 * torch.randn(n_rows, 2, generator=gen)  # every rank builds the same global batch
    a, b = shard_range(n_rows, world, rank)
    local = flows_cpu.log_prob(sd, specs, x[a:b])
    full = gather_rows(local, n_rows)
    ref = flows_cpu.log_prob(sd, specs, x)
    ok = torch.allclose(full, ref, rtol=1e-6, atol=1e-6)  # CPU kernels vectorise differently per batch size
    # MC reduction: 6 samples of 5 images split 4 + 2 over the ranks
    lp = torch.log_softmax(torch.randn(6, 5, 10, generator=gen), -1)
    mine = lp[:4] if rank == 0 else lp[4:]
    probs = reduce_mc_probs(mine.reshape(-1, 10), 5)
    ok = ok and torch.allclose(probs, lp.exp().mean(0), atol=1e-6)
    # ActNorm data-dependent init: rank 0 initialises (stood in by a function that writes s, t), everyone ends up equal
    import torch_mnf.flows as nf
    from torch_mnf.distributed import sync_actnorm_init

    torch.manual_seed(100 + rank)  # different random init per rank on purpose
    model = nf.NormalizingFlow([nf.ActNormFlow(3), nf.Glow(3), nf.ActNormFlow(3)])

    def fake_init(m):
        for i, f in enumerate(m.flows):
            if hasattr(f, "data_dep_init_done"):
                f.s.data.fill_(0.25 * (i + 1))
                f.t.data.fill_(-1.5 * (i + 1))

    n_sync = sync_actnorm_init(model, init_fn=fake_init)
    ok = ok and n_sync == 2 and all(f.data_dep_init_done for f in model.flows if hasattr(f, "data_dep_init_done"))
    ok = ok and bool((model.flows[0].s == 0.25).all()) and bool((model.flows[2].t == -4.5).all())
    torch.save(ok, os.path.join(out_dir, f"ok{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [64, 77])
def test_sharded_log_prob_gather(tmp_path, n_rows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(tmp_path / f"ok{r}.pt") for r in range(world))
