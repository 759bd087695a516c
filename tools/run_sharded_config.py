"""Whole-job run of the multi-GPU BASELINE configs under torchrun (one rank per GPU), strong scaling:

    torchrun --nproc-per-node 8 tools/run_sharded_config.py cfg3     # ZDT2 d=10, N=4096, 16 M candidates
    torchrun --nproc-per-node 8 tools/run_sharded_config.py cfg4     # DTLZ2 d=8, N=2048, m=3, 8 M candidates + Pareto
    torchrun --nproc-per-node 8 tools/run_sharded_config.py hl int8  # N=4096, d=6, 16 M candidates, INT8 engine

Candidates are generated on the device per shard from a generator seeded by the global chunk index, so any rank
count produces the same global candidate set.  Timing: CUDA events per rank, MAX over ranks."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesopt_smart_b200 import distributed as bd  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from bayesopt_smart_b200.pareto import _mask_against, pareto_mask_device  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)

CONFIGS = {"cfg3": dict(fn="zdt2", n=4096, d=10, m=2, ls=0.5, total=16_000_000, pareto=False),
           "cfg4": dict(fn="dtlz2", n=2048, d=8, m=3, ls=0.5, total=8_000_000, pareto=True),
           "cfg2x": dict(fn="zdt1", n=1024, d=6, m=2, ls=0.3, total=8_000_000, pareto=False),
           # the north star's headline shape: N=4096, d=6, 2 objectives
           "hl": dict(fn="zdt1", n=4096, d=6, m=2, ls=0.3, total=16_000_000, pareto=False)}
CHUNK = 250_000  # candidates per generator chunk (global chunk index = seed)


def shard_candidates(lo, hi, d, dev):
    parts = []
    for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
        g = torch.Generator(device=dev).manual_seed(1234 + c)
        block = torch.rand(CHUNK, d, dtype=torch.float64, device=dev, generator=g)
        a, b = max(lo, c * CHUNK) - c * CHUNK, min(hi, (c + 1) * CHUNK) - c * CHUNK
        parts.append(block[a:b])
    return torch.cat(parts).contiguous()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    engine = sys.argv[2] if len(sys.argv) > 2 else "dmma"  # variance engine: dmma | int8
    cfg = CONFIGS[tag]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    n, d, m = cfg["n"], cfg["d"], cfg["m"]
    x, y, mu0, var0 = orc.make_training_set(cfg["fn"], n, d, seed=0)
    ls, betas = np.full(m, cfg["ls"]), np.full(m, 2.0)
    lo, hi = bd.shard_range(cfg["total"], world, rank)
    cand = shard_candidates(lo, hi, d, dev)
    gp = DeviceGP(dev, variance_engine=engine)
    xd, yd = to_device(x, device=dev), to_device(y, device=dev)
    want = ("acq", "ucb") if cfg["pareto"] else ("acq",)
    out = {k: torch.empty((hi - lo,) if k == "acq" else (m, hi - lo), dtype=torch.float64, device=dev) for k in want}

    def step():
        gp.fit(xd, yd, mu0, var0, ls, n)
        gp.score(cand, betas, want=want, out=out)
        vals, idx = bd.select_next_batch_sharded(gp, cand, out["acq"], xd, 3, lo)
        front = None
        if cfg["pareto"]:
            rows = out["ucb"].T.contiguous()
            mask = bd.pareto_mask_sharded(rows, pareto_mask_device, _mask_against)
            front = int(mask.sum().item())
        return vals, idx, front

    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    times = []
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vals, idx, front = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    fronts = torch.tensor([front or 0], device=dev)
    if world > 1:
        dist.all_reduce(fronts)
    if rank == 0:
        best = min(times)
        flops = float(cfg["total"]) * m * n * n
        print(json.dumps({"config": tag, "variance_engine": engine, "n_gpus": world, "n_train": n, "dims": d, "objectives": m,
                          "candidates_total": cfg["total"], "step_s": best, "cand_per_s": cfg["total"] / best,
                          "algorithmic_tflops_total": flops / best / 1e12,
                          "algorithmic_tflops_per_gpu": flops / best / 1e12 / world,
                          "batch_idx": idx.cpu().tolist(), "batch_val": vals.cpu().tolist(),
                          "pareto_front_of_ucb_vectors": int(fronts.item()) if cfg["pareto"] else None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
