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
    if t.is_cuda and dist.get_backend(group) == "gloo":
        # gloo has no CUDA all-gather: stage through the host (used when several ranks share one GPU in tests; the
        # lists exchanged here are a few hundred bytes)
        host = t.cpu()
        parts = [torch.empty_like(host) for _ in range(world)]
        dist.all_gather(parts, host, group=group)
        return torch.cat(parts, dim=0).to(t.device)
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


def _local_list(gp, cand_shard, acq_shard, evaluated, k: int, index_base: int):
    """This rank's best ``k`` (value, global index) pairs, exactly ``k`` long: evaluated rows carry index -1
    ("no entry" for the merge), short shards are padded with empty entries."""
    dev = acq_shard.device
    n = acq_shard.numel()
    kk = min(n, k)
    if kk > 0:
        vals, idx, flags = gp.select_listed(cand_shard, acq_shard, evaluated, kk, index_base)
        idx = torch.where(flags.bool(), torch.full_like(idx, -1), idx)
    else:
        vals = torch.empty(0, dtype=torch.float64, device=dev)
        idx = torch.empty(0, dtype=torch.int64, device=dev)
    got = vals.numel()  # the kernel's own clamp (BO_MAX_TOPK) is honoured by padding from the RETURNED length
    if got < k:
        vals = torch.cat([vals, torch.full((k - got,), float("nan"), dtype=vals.dtype, device=dev)])
        idx = torch.cat([idx, torch.full((k - got,), -1, dtype=idx.dtype, device=dev)])
    return vals, idx


def select_next_batch_sharded(gp, cand_shard: torch.Tensor, acq_shard: torch.Tensor, evaluated: torch.Tensor,
                              batch_size: int, index_base: int, group=None, slack: int = 16, max_list: int = None):
    """Distributed select_next_batch (reference acquisition.py:116-144 over the union of all shards).

    Each rank lists its best ``k = batch_size + slack`` candidates (``k`` is clamped to BO_MAX_TOPK so that every
    rank sends a list of the same length), marks rows equal to an evaluated point, all ranks exchange the lists
    and merge them with the same comparator.  If fewer than ``batch_size`` valid entries survive while some rank
    could still list more, ``k`` grows (x4, up to the cap) and the exchange is repeated -- every rank sees the same
    merged list, so all take the same decision.  At the cap the evaluated candidates are masked out exhaustively
    (``DeviceGP.mask_evaluated``) before listing, which makes the result exact for any number of evaluated points.
    Returns (values, global indices) tensors of length ``batch_size`` on the device, identical on every rank;
    exhausted candidate sets give trailing entries with index -1.
    """
    from . import _lib  # BO_MAX_TOPK

    cap = int(max_list or _lib.BO_MAX_TOPK)
    dev = acq_shard.device
    n_local = torch.tensor([acq_shard.numel()], dtype=torch.int64, device=dev)
    n_max = int(all_gather_cat(n_local, group).max().item())  # longest shard: no rank can list more than this
    k = max(1, min(batch_size + max(int(slack), 0), cap))
    masked = False
    while True:
        vals, idx = _local_list(gp, cand_shard, acq_shard, evaluated, k, index_base)
        gv, gi = gather_topk(vals, idx, group)
        mv, mi = merge_topk(gp, gv, gi, batch_size)
        if mi.numel() < batch_size:  # fewer pairs than the batch in total (tiny candidate sets)
            pad = batch_size - mi.numel()
            mv = torch.cat([mv, torch.full((pad,), float("nan"), dtype=mv.dtype, device=dev)])
            mi = torch.cat([mi, torch.full((pad,), -1, dtype=mi.dtype, device=dev)])
        enough = bool((mi >= 0).all().item())
        exhausted = k >= min(n_max, cap)
        if enough or (exhausted and (masked or k >= n_max)):
            return mv, mi
        if exhausted:
            acq_shard = gp.mask_evaluated(acq_shard, cand_shard, evaluated) if acq_shard.numel() else acq_shard
            masked = True
            k = max(1, min(batch_size + 16, cap))
            continue
        k = min(cap, k * 4)


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
