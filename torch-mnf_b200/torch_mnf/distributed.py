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
