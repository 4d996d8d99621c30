"""Utterance sharding across ranks (one process per GPU).

Utterances are independent units, so the upstream forward needs no data-path collective: rank r runs the path on
its contiguous slice of the batch. Two pieces of cross-rank state keep the result identical to the un-sharded
batch: the padded length ``Lmax`` (it fixes the padding, the layer-0 GroupNorm statistics and the frame-mask rule)
and, at the end, ONE all-gather of the Featurizer output ``[B/G, T, D]`` (NCCL over NVLink on the GPU box; the
same code runs over gloo on CPU for the tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n_items`` for ``rank`` (first ``n_items % world`` ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(items: Sequence, rank: int, world: int) -> List:
    lo, hi = shard_bounds(len(items), rank, world)
    return list(items[lo:hi])


def global_max_len(local_lens: Sequence[int], device=None) -> int:
    """max over all ranks of the utterance lengths (1 scalar all-reduce); local max when not distributed."""
    m = max(local_lens) if len(local_lens) else 0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([m], dtype=torch.int64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m = int(t.item())
    return m


def gather_features(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather the per-rank feature blocks ``[n_local, T, D]`` into ``[n_items, T, D]`` (rank-major = original
    order, because shards are contiguous). Equal shards use a single ``all_gather_into_tensor``; uneven shards pad
    to the largest shard first (still one collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    biggest = max(counts)
    if local.shape[0] != biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    if all(c == biggest for c in counts):
        return out
    parts = [out[r * biggest : r * biggest + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
