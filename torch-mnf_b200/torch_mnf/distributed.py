"""Sharding helpers for the data-parallel axes of the hot path.

Points (flows) and Monte-Carlo rows (MNF layers) are independent given the weights, so ranks own
contiguous row blocks, weights are replicated, and the only collective is the final gather /
reduction of results over NCCL.  Noise is keyed by the GLOBAL row index (``row_offset``), so a
sharded run reproduces the single-device numbers.  Everything here is backend-agnostic
(``nccl`` on GPUs, ``gloo`` in the CPU tests)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of rank; the first n_rows % world ranks get one extra row."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """All-gather per-row results (log-probs, log-dets, class log-probabilities) of a batch sharded
    with ``shard_range``.  Equal shards use one all_gather_into_tensor; ragged ones pad to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_rows, world, r) for r in range(world)]
    width = local.shape[1:]
    biggest = max(b - a for a, b in sizes)
    if local.size(0) != sizes[rank][1] - sizes[rank][0]:
        raise ValueError("local shard has the wrong number of rows")
    if all(b - a == biggest for a, b in sizes):
        out = local.new_empty((n_rows, *width))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((biggest, *width))
    padded[: local.size(0)] = local
    buf = local.new_empty((world * biggest, *width))
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * biggest: r * biggest + (b - a)] for r, (a, b) in enumerate(sizes)])


def reduce_mc_probs(log_probs: torch.Tensor, n_images: int, group=None) -> torch.Tensor:
    """MC predictive mean over ALL ranks' samples: ``log_probs`` [S_local * n_images, C] (sample-major,
    as produced by forward(x.repeat(S, ...))) -> [n_images, C] mean class probabilities.  Only the
    [n_images, C] sums cross the wire (all_reduce), never the raw samples."""
    s_local = log_probs.size(0) // n_images
    sums = log_probs.exp().view(s_local, n_images, -1).sum(0)
    count = torch.tensor([float(s_local)], device=log_probs.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, group=group)
        dist.all_reduce(count, group=group)
    return sums / count


def sync_actnorm_init(model, x_init=None, group=None, src: int = 0, init_fn=None) -> int:
    """Data-dependent ActNorm initialisation for a model replicated over ranks: the statistics must
    be those of ONE batch, not of each rank's shard.  Rank ``src`` runs the first ``inverse`` on ``x_init`` (which
    initialises every pending ActNormFlow from its own input, affine_constant_flow.py:42-50), then ``s`` and ``t`` of
    all ActNorm flows are broadcast and marked initialised everywhere.  ``init_fn(model)`` replaces the default
    ``model.inverse(x_init)`` (used by the CPU tests).  Returns the number of flows synchronised."""
    flows = [f for f in model.flows if hasattr(f, "data_dep_init_done")]
    if not flows:
        return 0
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else src
    if rank == src and any(not f.data_dep_init_done for f in flows):
        with torch.no_grad():
            if init_fn is not None:
                init_fn(model)
            else:
                model.inverse(x_init)
    for f in flows:
        if multi:
            for t in (f.s, f.t):
                dist.broadcast(t.data, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
        f.data_dep_init_done = True
        f.__dict__["_program_salt"] = f.__dict__.get("_program_salt", 0) + 1  # cached flow programs re-pack s, t
    return len(flows)


class PeerGather:
    """Gather buffer in NVLink peer memory (torch symmetric memory) for per-row results of a sharded batch.

    ``buffer`` is a ``[world * n_rows]`` fp32 tensor that exists at the same offset on every rank.  The flow kernel
    stores its log-probs into the local slice AND -- through ``gather_out()`` -- into every peer's copy while it
    computes (NVLS multicast when the fabric supports it, else one store per peer), so no collective moves data
    afterwards; ``barrier()`` orders the step against the peers' reads."""

    def __init__(self, n_rows: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        group = group or dist.group.WORLD
        self.world, self.rank, self.n_rows = dist.get_world_size(group), dist.get_rank(group), n_rows
        if self.world - 1 > 8:
            raise ValueError("PeerGather supports up to 9 ranks per node")
        self.buffer = symm.empty(self.world * n_rows, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buffer, group)
        self._lib = _lib
        self.multicast = bool(getattr(self.handle, "has_multicast_support", False)) and int(self.handle.multicast_ptr) != 0

    def local_slice(self) -> torch.Tensor:
        return self.buffer[self.rank * self.n_rows:(self.rank + 1) * self.n_rows]

    def gather_out(self, use_multicast: bool = True):
        g = self._lib.GatherOut()
        g.row_offset = self.rank * self.n_rows
        if use_multicast and self.multicast:
            g.n_peers, g.multicast_ptr = 0, int(self.handle.multicast_ptr)
        else:
            peers = [int(p) for r, p in enumerate(self.handle.buffer_ptrs) if r != self.rank]
            g.n_peers = len(peers)
            for i, p in enumerate(peers):
                g.peer_ptrs[i] = p
        return g

    def barrier(self):
        self.handle.barrier()
