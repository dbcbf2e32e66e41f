"""Multi-GPU plumbing of the feature extraction.

Event clips are independent, so the path shards by *video* (5 clips each, contiguous
blocks so rank-order concatenation keeps dataset order -- the contiguous analogue of
the reference's DistributedSampler(shuffle=False), utils/dat_utils.py:25-33,59) with no
collective on the forward path.  The only exchange is one all-gather of the
[n_local_clips, D] fp32 features at the end (the reference does this through pickle
files, vidsitu_code/evl_vsitu.py:99-115, or not at all, feat_extractor.py:123).
One process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) block of `n_items` owned by `rank`; the first n_items % world
    ranks own one extra item."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


_GATHER_BUFS: dict = {}


def _gather_bufs(local: torch.Tensor, longest: int, world: int):
    """Padded send / receive buffers of gather_rows, allocated once per (shape, dtype, device, world)."""
    key = (tuple(local.shape[1:]), local.dtype, str(local.device), longest, world)
    bufs = _GATHER_BUFS.get(key)
    if bufs is None:
        padded = local.new_zeros((longest,) + tuple(local.shape[1:]))
        out = local.new_empty((world * longest,) + tuple(local.shape[1:]))
        bufs = _GATHER_BUFS[key] = (padded, out)
    return bufs


def gather_rows(local: torch.Tensor, n_total_rows: int, rows_per_item: int = 5) -> torch.Tensor:
    """All-gather row blocks of unequal length: every rank contributes the rows of its
    shard_range() of items (rows_per_item rows each); returns the [n_total_rows, D] tensor in
    item order on every rank.  Shards are padded to the longest one so a single
    all_gather_into_tensor (NCCL) / all_gather (gloo) moves everything.  On NCCL the buffers are allocated once and
    re-used: the returned tensor is valid until the next call with the same shapes."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n_items = n_total_rows // rows_per_item
    counts = [c * rows_per_item for c in shard_counts(n_items, world)]
    if local.shape[0] != counts[dist.get_rank()]:
        raise ValueError(f"rank {dist.get_rank()} holds {local.shape[0]} rows, expected {counts[dist.get_rank()]}")
    longest = max(counts)
    equal = all(c == longest for c in counts)
    if dist.get_backend() == "nccl":
        padded, out = _gather_bufs(local, longest, world)
        if equal and local.is_contiguous():
            dist.all_gather_into_tensor(out, local)      # equal shards: no staging copy, the result is `out` itself
            return out
        padded[: local.shape[0]] = local
        dist.all_gather_into_tensor(out, padded)
        parts = list(out.split(longest, dim=0))
    else:
        padded = local.new_zeros((longest,) + tuple(local.shape[1:]))
        padded[: local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
