"""World-size-2 gloo tests (CPU) of the host-side sharding / exchange logic in
bayesopt_smart_b200/distributed.py.  The merge comparator and the dominance test are supplied by the
oracle here (the product supplies CUDA kernels); what is under test is partitioning, the all-gathers and
the index bookkeeping."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesopt_smart_b200 import distributed as bd
from oracle import gp_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _HostMerge:
    """Stand-in for DeviceGP's CUDA top-k kernels with the same total order, on CPU tensors."""

    def __init__(self, max_topk=1024):
        self.max_topk = max_topk  # the CUDA kernel clamps k to BO_MAX_TOPK silently; so does this stand-in

    def topk(self, acq, k, index_base=0):
        a = acq.numpy()
        order = orc.ranked_indices(a)[: min(k, self.max_topk)]
        return torch.from_numpy(a[order].copy()), torch.from_numpy(order + index_base)

    def select_listed(self, cand, acq, evaluated, k, index_base=0):
        vals, idx = self.topk(acq, k, index_base)
        return vals, idx, self.match_rows(idx, cand, evaluated, index_base)

    def mask_evaluated(self, acq, cand, evaluated):
        hit = (cand[:, None, :] == evaluated[None, :, :]).all(dim=2).any(dim=1)
        return torch.where(hit, torch.full_like(acq, float("nan")), acq)

    def match_rows(self, idx, cand, evaluated, index_base=0):
        rows = cand[idx - index_base]
        hit = (rows[:, None, :] == evaluated[None, :, :]).all(dim=2).any(dim=1)
        return hit.to(torch.uint8)

    def topk_merge(self, vals, idx, k):
        v, i = vals.numpy(), idx.numpy()
        live = i >= 0  # index -1 marks "no entry" (same convention as the CUDA merge kernel)
        v, i = v[live], i[live]
        key = np.where(np.isnan(v), -np.inf, v)
        order = np.lexsort((i, -key))[:k]
        return torch.from_numpy(v[order].copy()), torch.from_numpy(i[order].copy())


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        n_cand, d = 1001, 3
        cand = rng.integers(0, 50, size=(n_cand, d)).astype(np.float64)
        acq = rng.normal(size=n_cand)
        acq[10] = acq[700] = acq.max() + 1.0  # cross-shard tie at the top
        evaluated = cand[[int(np.argsort(-acq)[2]), 5]]
        lo, hi = bd.shard_range(n_cand, world, rank)
        gp = _HostMerge()
        vals, idx = bd.select_next_batch_sharded(gp, torch.from_numpy(cand[lo:hi]), torch.from_numpy(acq[lo:hi]),
                                                 torch.from_numpy(evaluated), 4, lo)
        want_rows, want_idx = orc.ref_select_next_batch(cand, acq, evaluated, 4)
        ok_sel = np.array_equal(idx.numpy(), want_idx)

        y = rng.normal(size=(600, 3))
        y[7] = y[450]
        plo, phi = bd.shard_range(600, world, rank)

        def local_mask(t):
            return torch.from_numpy(orc.pareto_mask_definition(t.numpy()).astype(np.uint8))

        def against(a, z):
            an, zn = a.numpy(), z.numpy()
            ge = np.all(zn[None, :, :] >= an[:, None, :], axis=2)
            gt = np.any(zn[None, :, :] > an[:, None, :], axis=2)
            return torch.from_numpy((~np.any(ge & gt, axis=1)).astype(np.uint8))

        mask = bd.pareto_mask_sharded(torch.from_numpy(y[plo:phi]), local_mask, against)
        ok_par = np.array_equal(mask.numpy().astype(bool), orc.pareto_mask_definition(y)[plo:phi])
        ret[rank] = (ok_sel, ok_par, idx.tolist())
    finally:
        dist.destroy_process_group()


def _worker_large_batch(rank, world, port, ret):
    """batch + slack beyond the kernel's list cap, a short last shard, and more evaluated rows
    at the top of the ranking than any list can hold -- lists must stay equally long on every rank, the list must
    grow, and the exhaustive mask must take over at the cap."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        n_cand, d, cap = 61, 2, 16          # shards: 31 + 30 rows; with world 2 and cap 16 < batch + slack
        cand = np.stack([np.arange(n_cand), rng.integers(0, 9, n_cand)], axis=1).astype(np.float64)
        acq = rng.normal(size=n_cand)
        order = np.argsort(-acq)
        results = {}
        for name, n_eval, batch, slack in [("grow", 10, 5, 0), ("mask", 40, 6, 1000), ("exhausted", 58, 6, 16)]:
            evaluated = cand[order[:n_eval]]  # the best n_eval candidates were all evaluated already
            lo, hi = bd.shard_range(n_cand, world, rank)
            gp = _HostMerge(max_topk=cap)
            vals, idx = bd.select_next_batch_sharded(gp, torch.from_numpy(cand[lo:hi]), torch.from_numpy(acq[lo:hi]),
                                                     torch.from_numpy(evaluated), batch, lo, slack=slack, max_list=cap)
            _, want_idx = orc.ref_select_next_batch(cand, acq, evaluated, batch)
            got = idx.numpy()
            results[name] = (got[got >= 0].tolist(), list(want_idx), int(idx.numel()))
        # a shard shorter than the list: 3 candidates over 2 ranks (2 + 1), batch 2
        lo, hi = bd.shard_range(3, world, rank)
        tiny_c, tiny_a = cand[:3], np.array([0.5, 2.0, 1.0])
        vals, idx = bd.select_next_batch_sharded(_HostMerge(max_topk=cap), torch.from_numpy(tiny_c[lo:hi]),
                                                 torch.from_numpy(tiny_a[lo:hi]), torch.from_numpy(cand[50:51]), 2, lo,
                                                 max_list=cap)
        results["tiny"] = (idx.tolist(), [1, 2], int(idx.numel()))
        ret[rank] = results
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_selection_large_batch_short_shard_many_evaluated():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_large_batch, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            for name, (got, want, length) in ret[r].items():
                assert got == want, (r, name, got, want)
            assert ret[r]["grow"][2] == 5 and ret[r]["mask"][2] == 6 and ret[r]["exhausted"][2] == 6
            assert len(ret[r]["exhausted"][0]) == 3  # only 3 un-evaluated candidates exist
        assert ret[0] == ret[1]


def test_shard_range_partition():
    for n, w in [(10, 3), (1000000, 8), (5, 8), (0, 2)]:
        spans = [bd.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo <= hi for lo, hi in spans)


@pytest.mark.timeout(120)
def test_two_rank_selection_and_pareto_exchange():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        assert all(ret[r][0] for r in range(world)), dict(ret)
        assert all(ret[r][1] for r in range(world)), dict(ret)
        assert ret[0][2] == ret[1][2]  # identical batch on every rank
