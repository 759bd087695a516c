"""Measure the BASELINE.json configs (per-GPU share) on one B200 and spot-check each against the CPU oracle.

    python tools/run_configs.py [cfg2 cfg3 cfg4 cfg5 pareto]      -> gpurun_out/configs.jsonl

Multi-GPU configs are run here as ONE rank's shard (M / n_gpus candidates): the path has no data-path
collective, so the per-rank time is the job time up to the tiny top-k exchange (bench.py --gpus N measures that).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesopt_smart_b200 import _lib  # noqa: E402
from bayesopt_smart_b200 import numba_kernels as nk  # noqa: E402
from bayesopt_smart_b200.acquisition import exact_hvi_device  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from bayesopt_smart_b200.pareto import pareto_mask_device  # noqa: E402
from oracle import gp_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "configs.jsonl"), "a")
EPS = np.finfo(np.float64).eps


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


def ev_time(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def gp_config(tag, fn, n, d, m, ls, n_cand, shards, check=1500):
    x, y, mu0, var0 = orc.make_training_set(fn, n, d, seed=0)
    lsv, betas = np.full(m, ls), np.full(m, 2.0)
    gp = DeviceGP()
    xd, yd = to_device(x), to_device(y)
    t_fit = ev_time(lambda: gp.fit(xd, yd, mu0, var0, lsv, n), warm=1, reps=2)
    g = torch.Generator(device="cuda").manual_seed(1)
    cand = torch.rand(n_cand, d, dtype=torch.float64, device="cuda", generator=g)
    out = {k: torch.empty((n_cand,) if k == "acq" else (m, n_cand), dtype=torch.float64, device="cuda")
           for k in ("mu", "var", "ucb", "acq")}
    t_score = ev_time(lambda: gp.score(cand, betas, out=out), warm=1, reps=2)
    t_top = ev_time(lambda: gp.topk(out["acq"], 19), warm=1, reps=3)
    # oracle spot check on a strided sample of the candidates
    sel = torch.arange(0, n_cand, max(1, n_cand // check), device="cuda")[:check]
    cs = cand[sel].cpu().numpy()
    t0 = time.perf_counter()
    fit = orc.chol_fit(x, y, mu0, var0, lsv, n)
    mu_o, var_o = orc.chol_predict(fit, x, cs, mu0, var0, lsv, n)
    t_oracle = time.perf_counter() - t0
    cond = max(np.linalg.cond(fit["kernel"][o] + 1e-6 * np.eye(n)) for o in range(m)) if n <= 2048 else None
    err_mu = max(np.abs(out["mu"][o][sel].cpu().numpy() - mu_o[o]).max() / np.sqrt(var0[o]) for o in range(m))
    err_var = max(np.abs(out["var"][o][sel].cpu().numpy() - var_o[o]).max() / var0[o] for o in range(m))
    flops = float(n_cand) * m * n * n
    emit(kind="gp", tag=tag, n=n, d=d, m=m, n_cand_this_gpu=n_cand, n_gpus_in_config=shards, fit_s=t_fit,
         score_s=t_score, topk_s=t_top, cand_per_s_per_gpu=n_cand / (t_fit + t_score + t_top),
         score_tflops_algorithmic=flops / t_score / 1e12, err_mu_std=err_mu, err_var_std=err_var, cond=cond,
         oracle_sample=int(sel.numel()), oracle_s=t_oracle)
    return gp, cand, out, x, y, mu0, var0


def main():
    which = set(sys.argv[1:]) or {"cfg2", "cfg3", "cfg4", "cfg5", "pareto"}
    emit(kind="device", name=torch.cuda.get_device_name(0))
    if "cfg2" in which:
        gp_config("cfg2_zdt1_n1024_d6_m2_1M", "zdt1", 1024, 6, 2, 0.3, 1_000_000, 1)
    if "cfg3" in which:
        gp_config("cfg3_zdt2_n4096_d10_m2_16M_over_8", "zdt2", 4096, 10, 2, 0.5, 2_000_000, 8)
    if "cfg4" in which:
        gp, cand, out, x, y, mu0, var0 = gp_config("cfg4_dtlz2_n2048_d8_m3_8M_over_8", "dtlz2", 2048, 8, 3, 0.5,
                                                   1_000_000, 8)
        # 3-objective exact HVI of the UCB vectors against the standardised training front + Pareto filter
        y_std = (y - mu0) / np.sqrt(var0)
        front = y_std[orc.pareto_mask_definition(y_std)]
        ref = y_std.min(axis=0) - 0.1
        t_hvi = ev_time(lambda: exact_hvi_device(out["ucb"], front, ref), warm=1, reps=2)
        hv = exact_hvi_device(out["ucb"], front, ref)
        sel = torch.arange(0, cand.shape[0], cand.shape[0] // 300, device="cuda")[:300]
        want = orc.exact_hvi(out["ucb"][:, sel].T.cpu().numpy(), front, ref)
        err = float(np.abs(hv[sel].cpu().numpy() - want).max() / max(1e-300, np.abs(want).max()))
        ucb_rows = out["ucb"].T.contiguous()
        t_par = ev_time(lambda: pareto_mask_device(ucb_rows), warm=1, reps=2)
        mask = pareto_mask_device(ucb_rows)
        emit(kind="hvi3_pareto", front_size=int(front.shape[0]), hvi_s=t_hvi, hvi_rel_err_vs_oracle=err,
             hvi_cand_per_s=cand.shape[0] / t_hvi, pareto_s=t_par, pareto_front_of_ucb=int(mask.sum().item()),
             n_points=int(ucb_rows.shape[0]))
    if "pareto" in which:
        g = torch.Generator(device="cuda").manual_seed(3)
        for n, m in [(8_000_000, 3), (8_000_000, 2)]:
            yv = torch.randn(n, m, dtype=torch.float64, device="cuda", generator=g)
            t = ev_time(lambda: pareto_mask_device(yv), warm=1, reps=2)
            mask = pareto_mask_device(yv).bool()
            front = yv[mask].cpu().numpy()
            ok_front = bool(orc.pareto_mask_definition(front).all())
            drop = yv[~mask][:: max(1, int((~mask).sum().item()) // 3000)].cpu().numpy()
            ge = np.all(front[None, :, :] >= drop[:, None, :], axis=2)
            gt = np.any(front[None, :, :] > drop[:, None, :], axis=2)
            emit(kind="pareto", n=n, m=m, seconds=t, points_per_s=n / t, front=int(mask.sum().item()),
                 front_is_nondominated=ok_front, sampled_dropped_all_dominated=bool(np.all(np.any(ge & gt, axis=1))))
    if "cfg5" in which:
        n, d, m = 4096, 6, 2
        x, y, mu0, _ = orc.make_training_set("zdt1", n, d, seed=0)
        ls = np.repeat(np.logspace(-1, 0.5, 16), 16)
        jit = np.tile(np.logspace(-8, -2, 16), 16)
        xd, yd = to_device(x), to_device(y)
        t0 = time.perf_counter()
        vals = nk.mll_batched(xd, yd, mu0, np.stack([ls, ls], axis=1), jit, n)
        torch.cuda.synchronize()
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        vals = nk.mll_batched(xd, yd, mu0, np.stack([ls, ls], axis=1), jit, n)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        # settings with jitter 1e-8 must equal compute_mll: check two of them against the oracle
        errs = []
        t_or = 0.0
        for s in (0, 8 * 16):
            t1 = time.perf_counter()
            want = orc.ref_compute_mll(x, y, np.zeros((m, n, n)), mu0, np.ones(m), np.full(m, ls[s]), n)
            t_or += time.perf_counter() - t1
            errs.append(abs(vals[s] - want) / abs(want))
        flops = 256 * m * (n**3 / 3.0)
        emit(kind="mll_sweep", tag="cfg5_256_settings_n4096_d6_m2", seconds=t, first_call_s=t_first,
             settings_per_s=256 / t, potrf_tflops=flops / t / 1e12, nan_settings=int(np.isnan(vals).sum()),
             rel_err_vs_oracle=errs, oracle_s_per_setting=t_or / 2)


if __name__ == "__main__":
    main()
