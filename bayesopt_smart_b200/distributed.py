"""Candidate-sharded data parallelism (one process per GPU, torch.distributed).

Every candidate is independent through K*, mean, variance, UCB and sum-UCB, so ranks score disjoint
contiguous blocks of the candidate index space with NO data-path collective; the factor (L, W, alpha)
is recomputed identically on every rank from the replicated training set.  Two small exchanges exist:

* batch selection: all-gather of each rank's top-k (value, global index) pairs, then the same
  deterministic merge on every rank (value desc, index asc) -> identical ``x_next`` everywhere;
* Pareto filtering: all-gather of the local fronts (size-prefixed, padded), dominance pass of the local
  front against the union.

The reference is single-process (no counterpart); SURVEY 8(e) is the contract.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_cand: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``: rank r owns [r*ceil(M/G), (r+1)*ceil(M/G))."""
    per = -(-int(n_cand) // int(world))
    lo = min(rank * per, n_cand)
    return lo, min(lo + per, n_cand)


def all_gather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equally-sized 1-D/2-D tensors from all ranks along dim 0 (NCCL or gloo)."""
    _, world = world_info(group)
    if world == 1:
        return t
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t, group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        out = torch.cat(parts, dim=0)
    return out


def gather_topk(vals: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks' (value, global index) lists, concatenated in rank order."""
    return all_gather_cat(vals, group), all_gather_cat(idx, group)


def merge_topk(gp, vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k of the gathered pairs with the device comparator (value desc, index asc)."""
    return gp.topk_merge(vals, idx, k)


def select_next_batch_sharded(gp, cand_shard: torch.Tensor, acq_shard: torch.Tensor, evaluated: torch.Tensor,
                              batch_size: int, index_base: int, group=None, slack: int = 16):
    """Distributed select_next_batch (reference acquisition.py:116-144 over the union of all shards).

    Each rank lists its best ``batch_size + slack`` candidates, masks rows equal to an evaluated point
    (value -> -inf), all ranks exchange the lists and merge.  Returns (values, global indices) tensors of
    length ``batch_size`` on the device, identical on every rank.
    """
    n = acq_shard.numel()
    _, world = world_info(group)
    k = batch_size + slack  # same list length on every rank (all-gather needs equal sizes)
    kk = min(n, k)
    vals, idx = gp.topk(acq_shard, kk, index_base) if kk > 0 else (
        torch.empty(0, dtype=torch.float64, device=acq_shard.device),
        torch.empty(0, dtype=torch.int64, device=acq_shard.device))
    if kk > 0:
        flags = gp.match_rows(idx, cand_shard, evaluated, index_base)
        # an evaluated row is dropped from the ranking entirely: index -1 marks "no entry" for the merge
        idx = torch.where(flags.bool(), torch.full_like(idx, -1), idx)
    if kk < k:  # short shard: pad the list with empty entries
        vals = torch.cat([vals, torch.full((k - kk,), float("nan"), dtype=vals.dtype, device=vals.device)])
        idx = torch.cat([idx, torch.full((k - kk,), -1, dtype=idx.dtype, device=idx.device)])
    gv, gi = gather_topk(vals, idx, group)
    return merge_topk(gp, gv, gi, batch_size)


def gather_ragged_rows(rows: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather (n_r, m) row blocks of different lengths: sizes first, then padded blocks."""
    _, world = world_info(group)
    if world == 1:
        return rows, [rows.shape[0]]
    n_local = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = all_gather_cat(n_local, group).tolist()
    width = max(max(sizes), 1)
    pad = torch.zeros((width, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    blocks = all_gather_cat(pad, group).reshape(world, width, rows.shape[1])
    return torch.cat([blocks[r, : sizes[r]] for r in range(world)], dim=0), sizes


def pareto_mask_sharded(y_shard: torch.Tensor, local_mask_fn: Callable[[torch.Tensor], torch.Tensor],
                        against_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], group=None) -> torch.Tensor:
    """Non-dominated mask of this rank's rows with respect to the union of all ranks' rows.

    ``local_mask_fn(y)`` -> uint8 mask within a set; ``against_fn(y, z)`` -> uint8 mask of y's rows not
    dominated by any row of z.  A row dominated by anything is dominated by a member of some local front
    (transitivity), so comparing the local front with the union of local fronts is exact.
    """
    local = local_mask_fn(y_shard).bool()
    front = y_shard[local].contiguous()
    union, _ = gather_ragged_rows(front, group)
    keep = against_fn(front, union.contiguous()).bool()
    mask = torch.zeros(y_shard.shape[0], dtype=torch.uint8, device=y_shard.device)
    pos = torch.nonzero(local, as_tuple=False).reshape(-1)
    mask[pos[keep]] = 1
    return mask
