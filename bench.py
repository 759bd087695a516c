#!/usr/bin/env python
"""bench.py -- candidate scores/sec of the acquisition hot path (GP predict + UCB + sum-UCB "HVI"
+ top-k batch selection) on N B200s, next to the reference's own CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

One "step" = one pass of the reference's per-iteration hot path (bayesian_optimization.py:129-207:
update_k, invert_k, update_k_star, update_mean, update_variance, standardize_objectives, update_ucb,
update_hypervolume_improvement, select_next_batch) over one synthetic candidate set.

HEADLINE (`value`, `e2e`, `roofline`): BASELINE.json configs[1], "cfg2": ZDT1, d = 6, N = 1024 training points,
2 objectives, 10^6-point candidate grid linspace(0,1,10)^6 per GPU, length scale 0.3, beta 2, batch 3, FP64 DMMA
variance engine.  N > 1: weak scaling -- every rank scores its own 10^6 candidates (rank 0 the grid, rank r a
counter-seeded uniform shard); the factor is recomputed per rank (deterministic, no broadcast) and the only
exchange is an all-gather of each rank's top-k (value, global index) pairs over NCCL, merged on device.
`int8_engine` repeats the same step with the INT8 tensor-core variance engine.

EXTRA BLOCKS in the same JSON line (`baseline_configs`, skipped with --no-extras): the other BASELINE.json
configs as whole jobs, STRONG-scaled over the N ranks, each with its own value / per-kernel rooflines and two
correctness assertions -- `sharded_topk_equals_gathered` (all-gather of the scores, single-shot top-k, compared
bit for bit with the merge of the per-rank lists) and `oracle_spot_check` (256 candidates per rank against the
CPU oracle; the oracle is used as the checker only):
    north_star  N=4096, d=6, m=2, 16 M device-generated candidates     (every N; both engines)
    cfg4        DTLZ2 d=8, N=2048, m=3, 8 M candidates + Pareto filter  (every N; both engines)
    cfg3        ZDT2 d=10, N=4096, m=2, 16 M candidates                 (N >= 2; both engines)
    cfg5        256-setting log-marginal-likelihood sweep at N=4096     (settings sharded over the ranks)
    hbm_passes  stand-alone score pass / top-k scan at 16 M candidates  (rank 0)
    incremental_fit  factor extended by a batch of 3 points vs rebuilt, N = 1024 / 4096 (rank 0)
    cfg1_loop   the reference's demo through BayesianOptimization(...).optimize(), Powell fit included (rank 0)

`cpu_baseline` / `--impl reference`: the UNMODIFIED reference functions (oracle/_ref, a git-ignored verbatim
copy made by oracle/make_ref.py, Numba JIT) on the host cores, candidates fed in chunks of 16 384; if Numba or
the copy is missing the NumPy port (oracle/gp_oracle.py) is timed instead and the reason is printed.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bayesopt_smart_b200.workloads import CONFIGS  # noqa: E402  (input definitions only, no hot-path arithmetic)

WORKLOAD = CONFIGS["cfg2"]
METRIC = "candidate scores/sec (GP predict+UCB+HVI)"
UNIT = "candidates/s"
EPS = float(np.finfo(np.float64).eps)
CPU_CHUNK = 16384


def headline_config(world: int) -> dict:
    """The `config` object of the JSON line -- identical for the GPU arm and the reference arm."""
    w = WORKLOAD
    return {"workload": w["name"], "n_train": w["n"], "dims": w["d"], "objectives": w["m"],
            "candidates_per_gpu": w["total"], "candidates_total": w["total"] * world, "batch_size": w["batch"],
            "length_scale": w["ls"], "beta": w["beta"],
            "step": "update_k + invert_k + update_k_star + update_mean + update_variance + standardize_objectives "
                    "+ update_ucb + update_hypervolume_improvement + select_next_batch "
                    "(bayesian_optimization.py:129-207)" + (" + all-gather/merge of per-rank top-k" if world > 1 else ""),
            "l2": "GPU arm: 256 MiB flush write between timed steps (untimed), K* staging per chunk 0.6 GB > L2; "
                  "CPU arm: working set per chunk 0.27 GB > LLC",
            "timing": "GPU arm: per-step CUDA events on the launching stream, max over ranks; CPU arm: perf_counter"}


# --------------------------------------------------------------------------------------- workload
def make_workload(rank: int = 0):
    from bayesopt_smart_b200.workloads import make_training_set

    w = WORKLOAD
    x, y, mu0, var0 = make_training_set(w["fn"], w["n"], w["d"], seed=0)
    if rank == 0:
        axes = [np.linspace(0.0, 1.0, w["grid_levels"])] * w["d"]
        cand = np.stack([g.ravel() for g in np.meshgrid(*axes, indexing="ij")], axis=-1)
    else:
        cand = np.random.default_rng(1000 + rank).random((w["grid_levels"] ** w["d"], w["d"]))
    ls = np.full(w["m"], w["ls"])
    betas = np.full(w["m"], w["beta"])
    return x, y, mu0, var0, np.ascontiguousarray(cand), ls, betas


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU reference arm
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:  # noqa: BLE001
        pass
    return cores


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


class CpuReference:
    """The reference's per-iteration hot path on the host cores.

    kind "reference": the unmodified functions of oracle/_ref (Numba JIT + OpenBLAS, exactly what a user of
    alebal123bal/BayesOpt_smart runs); kind "port": oracle/gp_oracle.py (NumPy/OpenBLAS restatement), used only
    when the copy or Numba is unavailable -- the reason is kept in `why_port` and printed with the result.
    """

    def __init__(self):
        cores = os.cpu_count() or 1
        os.environ["NUMBA_NUM_THREADS"] = str(cores)  # before numba is imported (torchrun sets OMP_NUM_THREADS=1)
        os.environ.pop("OMP_NUM_THREADS", None)
        self.why_port = None
        try:
            from oracle.make_ref import import_reference

            self.nk, self.aq, _ = import_reference()
            import numba

            self.kind = "reference"
            self.threads = int(numba.get_num_threads())
            self.impl = f"unmodified bayesopt/numba_kernels.py + acquisition.py (Numba {numba.__version__}, " \
                        f"{self.threads} Numba threads, OpenBLAS {blas_threads()} threads)"
        except Exception as exc:  # noqa: BLE001
            from oracle import gp_oracle as orc

            self.kind = "port"
            self.why_port = f"{type(exc).__name__}: {exc}"
            print(f"[bench] reference Numba path unavailable ({self.why_port}); timing the NumPy port instead",
                  file=sys.stderr)
            self.orc = orc
            self.threads = blas_threads()
            self.impl = "NumPy/OpenBLAS port of the reference functions (oracle/gp_oracle.py)"
        use_all_host_threads()
        self.x, self.y, self.mu0, self.var0, self.cand, self.ls, self.betas = make_workload(0)
        self.n, self.m = WORKLOAD["n"], WORKLOAD["m"]
        self.full_acq = np.random.default_rng(3).normal(size=self.cand.shape[0])

    # --- the nine functions, in the reference loop's order, on candidates [c0, c0 + count) in chunks
    def fit(self):
        n, m = self.n, self.m
        kmat = np.zeros((m, n, n))
        if self.kind == "reference":
            self.nk.update_k(kernel_matrix=kmat, x_vector=self.x, last_eval=0, current_eval=n,
                             prior_variance=self.var0, length_scales=self.ls)
            return self.nk.invert_k(current_eval=n, kernel_matrix=kmat)
        self.orc.ref_update_k(kmat, self.x, 0, n, self.var0, self.ls)
        return self.orc.ref_invert_k(n, kmat)

    def score_chunk(self, kinv, cc, acq_out):
        n, m = self.n, self.m
        ks = np.zeros((m, n, cc.shape[0]))
        mu = np.zeros((m, cc.shape[0]))
        var = np.zeros_like(mu)
        smu, svar, ucb = np.zeros_like(mu), np.zeros_like(mu), np.zeros_like(mu)
        if self.kind == "reference":
            nk, aq = self.nk, self.aq
            nk.update_k_star(k_star=ks, x_vector=self.x, input_space=cc, last_eval=0, current_eval=n,
                             prior_variance=self.var0, length_scales=self.ls)
            nk.update_mean(mu_objectives=mu, k_star=ks, inverted_kernel_matrix=kinv, y_vector=self.y,
                           prior_mean=self.mu0, current_eval=n)
            nk.update_variance(variance_objectives=var, k_star=ks, inverted_kernel_matrix=kinv,
                               prior_variance=self.var0, current_eval=n)
            nk.standardize_objectives(std_mu_objectives=smu, std_variance_objectives=svar, mu_objectives=mu,
                                      variance_objectives=var, prior_mean=self.mu0, prior_variance=self.var0)
            aq.update_ucb(ucb=ucb, mu_objectives=smu, variance_objectives=svar, betas=self.betas)
            aq.update_hypervolume_improvement(acquisition_values=acq_out, ucb=ucb)
        else:
            orc = self.orc
            orc.ref_update_k_star(ks, self.x, cc, 0, n, self.var0, self.ls)
            orc.ref_update_mean(mu, ks, kinv, self.y, self.mu0, n)
            orc.ref_update_variance(var, ks, kinv, self.var0, n)
            orc.ref_standardize_objectives(smu, svar, mu, var, self.mu0, self.var0)
            orc.ref_update_ucb(ucb, smu, svar, self.betas)
            orc.ref_update_hypervolume_improvement(acq_out, ucb)

    def select(self, acq):
        if self.kind == "reference":
            return self.aq.select_next_batch(input_space=self.cand, acquisition_values=acq,
                                             evaluated_points=self.x[: self.n], batch_size=WORKLOAD["batch"])
        return self.orc.ref_select_next_batch(self.cand, acq, self.x[: self.n], WORKLOAD["batch"])

    def sample_step(self, sample_cands: int) -> dict:
        """a1+a2 once, a3..a8 on `sample_cands` grid candidates (strided over the grid), a9 on the full-length
        score vector.  Returns the three wall times."""
        stride = max(1, self.cand.shape[0] // sample_cands)
        cand = np.ascontiguousarray(self.cand[::stride][:sample_cands])
        t0 = time.perf_counter()
        kinv = self.fit()
        t1 = time.perf_counter()
        acq = np.zeros(cand.shape[0])
        for c0 in range(0, cand.shape[0], CPU_CHUNK):
            self.score_chunk(kinv, cand[c0:c0 + CPU_CHUNK], acq[c0:c0 + CPU_CHUNK])
        t2 = time.perf_counter()
        self.select(self.full_acq)
        t3 = time.perf_counter()
        return {"fit_s": t1 - t0, "score_s": t2 - t1, "select_s": t3 - t2, "candidates": cand.shape[0]}

    def measure(self, sample_cands: int, steps: int, warmup: int) -> dict:
        for _ in range(max(1, warmup)):  # JIT compilation + first-touch
            self.sample_step(min(sample_cands, 2048))
        runs = [self.sample_step(sample_cands) for _ in range(steps)]
        total = WORKLOAD["total"]
        fit = float(np.mean([r["fit_s"] for r in runs]))
        per_cand = float(np.sum([r["score_s"] for r in runs]) / np.sum([r["candidates"] for r in runs]))
        select = float(np.mean([r["select_s"] for r in runs]))
        # a3..a8 are exactly linear in the candidate count and a1, a2, a9 run once per step, so the full 10^6-candidate
        # step costs fit + 10^6 * per_candidate + select; `value` is candidates/s of THAT step (the sample's own
        # rate would charge the fit 30x too often and flatter the GPU arm)
        full_step = fit + per_cand * total + select
        wall = float(np.sum([r["fit_s"] + r["score_s"] + r["select_s"] for r in runs]))
        return {"value": total / full_step, "extrapolated_full_step_s": full_step, "fit_s": fit,
                "per_candidate_s": per_cand, "select_full_vector_s": select, "wall_s": wall,
                "sample_candidates_per_step": runs[0]["candidates"], "steps": steps}

    def baseline_block(self, res: dict) -> dict:
        blk = {"value": res["value"], "unit": UNIT, "cores": self.threads, "kind": self.kind,
               "sample": f"{res['sample_candidates_per_step']} of 10^6 grid candidates per step x {res['steps']} "
                         f"steps in chunks of {CPU_CHUNK} (the reference materialises k_star (m, N, M)); a1+a2 and a9 "
                         "(on the full-length score vector) timed once per step; value = 10^6 / (fit + 10^6 x "
                         "measured seconds per candidate + select), a3..a8 being exactly linear in M",
               "implementation": self.impl, "fit_s": res["fit_s"], "per_candidate_us": 1e6 * res["per_candidate_s"],
               "select_full_vector_s": res["select_full_vector_s"],
               "extrapolated_full_step_s": res["extrapolated_full_step_s"], "host_cpus": os.cpu_count()}
        if self.why_port:
            blk["why_port"] = self.why_port
        return blk


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref = CpuReference()
    res = ref.measure(sample_cands=2 * CPU_CHUNK, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["wall_s"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(world),
        "cpu_baseline": ref.baseline_block(res),
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ms_per_step is the measured wall time of one bounded sample step; value is the throughput of the "
                "full 10^6-candidate step extrapolated from it (see cpu_baseline.sample)",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm helpers
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 6536.7, "fallback: 6536.7 GB/s (MEASURED_PEAKS.json absent; the figure the pool measured earlier)"


class Profile:
    """bo_profile_* (CUDA events recorded by the library around each of its launches, on the launching stream)."""

    KINDS = {"contraction": 0, "kstar": 1, "finalize": 2, "topk": 3, "fit": 4}

    def __init__(self, lib):
        self.lib = lib

    def start(self):
        self.lib.bo_profile_read_kernel(-1, None, None, None)
        self.lib.bo_profile_enable(1)

    def stop(self) -> dict:
        out = {}
        for name, kid in self.KINDS.items():
            ms, nl, wk = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
            self.lib.bo_profile_read_kernel(kid, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(wk))
            out[name] = (ms.value, int(nl.value), wk.value)
        self.lib.bo_profile_enable(0)
        return out


def kernel_rooflines(prof: dict, engine: str, step_s: float, peaks: dict) -> dict:
    """Per-kernel roofline objects from the library's own event timings of one (or more) steps."""
    names = {"dmma": ("trmm_sumsq_kernel", "kstar_pack_kernel"), "int8": ("oz_sumsq_kernel", "oz_kstar_digits_kernel")}
    out = {}
    ms, nl, wk = prof["contraction"]
    if ms > 0:
        tf = wk / (ms * 1e-3) / 1e12
        if engine == "dmma":
            out[names[engine][0]] = {"bound": "tensor", "achieved": tf, "peak": peaks["dgemm_tflops"],
                                     "unit": "TFLOP/s", "frac": tf / peaks["dgemm_tflops"] if peaks["dgemm_tflops"] else None,
                                     "launches": nl, "avg_launch_ms": ms / nl, "share_of_step": ms * 1e-3 / step_s}
        else:
            out[names[engine][0]] = {"bound": "tensor", "achieved": 21.0 * tf, "peak": peaks["i8_tops"],
                                     "unit": "TOP/s (int8 multiply + add)",
                                     "frac": 21.0 * tf / peaks["i8_tops"] if peaks["i8_tops"] else None,
                                     "fp64_equivalent_tflops": tf, "launches": nl, "avg_launch_ms": ms / nl,
                                     "share_of_step": ms * 1e-3 / step_s}
    for key, kname in (("kstar", names[engine][1]), ("finalize", "finalize_kernel"), ("topk", "topk kernels")):
        ms, nl, wk = prof[key]
        if ms > 0:
            gbs = wk / (ms * 1e-3) / 1e9
            out[kname] = {"bound": "hbm" if key != "kstar" else "fp64 pipe + hbm writes", "achieved": gbs,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "launches": nl,
                          "avg_launch_ms": ms / nl, "share_of_step": ms * 1e-3 / step_s}
    ms, nl, wk = prof["fit"]
    if ms > 0:
        out["fit (gram + blocked Cholesky + W = L^-1 + alpha + pack)"] = {
            "bound": "latency / tensor", "achieved": wk / (ms * 1e-3) / 1e12, "peak": peaks["dgemm_tflops"],
            "unit": "TFLOP/s", "frac": wk / (ms * 1e-3) / 1e12 / peaks["dgemm_tflops"] if peaks["dgemm_tflops"] else None,
            "launches": nl, "avg_launch_ms": ms / nl, "share_of_step": ms * 1e-3 / step_s}
    return out


def oracle_spot_check(out, cand_dev, x, y, mu0, var0, ls, n, m, count, cond, fit_cache, world, rank):
    """CHECKER (not measured, not shipped): `count` evenly spaced candidates of EVERY rank's shard against the CPU
    oracle's Cholesky-form prediction (oracle/gp_oracle.py chol_fit / chol_predict, pinned to the reference's golden
    vectors by tests/test_oracle_golden.py).  The samples (candidate rows + the GPU's mu / var for them) are gathered
    on rank 0, which factors once per config with all host threads -- eight ranks factoring a 4096 x 4096 system at
    the same time on 16 cores took minutes.  Tolerance in standardised units: max(1e-9, 10 eps cond)."""
    import torch
    import torch.distributed as dist

    from bayesopt_smart_b200 import distributed as bd

    tau = max(1e-9, 10 * EPS * cond)
    dev = cand_dev.device
    n_c, d = cand_dev.shape
    pack = torch.full((count, d + 2 * m + 1), float("nan"), dtype=torch.float64, device=dev)  # last column: valid flag
    if n_c > 0:
        k = min(count, n_c)
        sel = torch.linspace(0, n_c - 1, k, device=dev).long()
        pack[:k, :d] = cand_dev[sel]
        pack[:k, d:d + m] = out["mu"][:, sel].T
        pack[:k, d + m:d + 2 * m] = out["var"][:, sel].T
        pack[:k, -1] = 1.0
    allp = bd.all_gather_cat(pack) if world > 1 else pack
    res = torch.zeros(3, dtype=torch.float64, device=dev)  # [max mu err, max var err, checked candidates]
    if rank == 0:
        from oracle import gp_oracle as orc

        use_all_host_threads()
        rows = allp[allp[:, -1] == 1.0].cpu().numpy()
        if "fit" not in fit_cache:
            fit_cache["fit"] = orc.chol_fit(x, y, mu0, var0, ls, n)
        mu_o, var_o = orc.chol_predict(fit_cache["fit"], x, np.ascontiguousarray(rows[:, :d]), mu0, var0, ls, n)
        emu = max(float(np.abs(rows[:, d + o] - mu_o[o]).max() / np.sqrt(var0[o])) for o in range(m))
        evar = max(float(np.abs(rows[:, d + m + o] - var_o[o]).max() / var0[o]) for o in range(m))
        res = torch.tensor([emu, evar, float(rows.shape[0])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)  # the other ranks hold zeros: this is a broadcast from rank 0
    emu, evar, checked = (float(v) for v in res.tolist())
    return {"candidates_per_rank": count, "candidates_checked": int(checked), "max_abs_mu_err_standardised": emu,
            "max_abs_var_err_standardised": evar, "tolerance": tau, "within_tolerance": bool(emu <= tau and evar <= tau)}


def run_sharded_job(tag, engine, dev, world, rank, peaks, lib, fit_cache):
    """One BASELINE config as a whole job, strong-scaled: rank r scores candidates shard_range(total, world, r)."""
    import torch
    import torch.distributed as dist

    from bayesopt_smart_b200 import distributed as bd
    from bayesopt_smart_b200.engine import DeviceGP, to_device
    from bayesopt_smart_b200.pareto import _mask_against, pareto_mask_device
    from bayesopt_smart_b200.workloads import make_training_set, shard_candidates

    cfg = CONFIGS[tag]
    n, d, m, total, k = cfg["n"], cfg["d"], cfg["m"], cfg["total"], cfg["batch"]
    x, y, mu0, var0 = make_training_set(cfg["fn"], n, d, seed=0)
    ls, betas = np.full(m, cfg["ls"]), np.full(m, cfg["beta"])
    lo, hi = bd.shard_range(total, world, rank)
    cand = shard_candidates(lo, hi, d, dev)
    gp = DeviceGP(dev, variance_engine=engine)
    xd, yd = to_device(x, device=dev), to_device(y, device=dev)
    want = ("mu", "var", "acq") + (("ucb",) if cfg.get("pareto") else ())
    out = {key: torch.empty((hi - lo,) if key == "acq" else (m, hi - lo), dtype=torch.float64, device=dev)
           for key in want}

    def step(c, o):
        gp.fit(xd, yd, mu0, var0, ls, n)
        gp.score(c, betas, want=want, out=o)
        vals, idx = bd.select_next_batch_sharded(gp, c, o["acq"], xd, k, lo)
        front = None
        if cfg.get("pareto"):
            mask = bd.pareto_mask_sharded(o["ucb"].T.contiguous(), pareto_mask_device, _mask_against)
            front = mask.sum()
        return vals, idx, front

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: same launch sequence on the first 1/16 of the shard (every kernel, every collective)
    wn = max(1, (hi - lo) // 16)
    step(cand[:wn], {key: v[..., :wn] if v.dim() == 1 else v[:, :wn].contiguous() for key, v in out.items()})
    step(cand[:wn], {key: v[..., :wn] if v.dim() == 1 else v[:, :wn].contiguous() for key, v in out.items()})
    prof = Profile(lib)
    sync()
    prof.start()
    lib.bo_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vals, idx, front = step(cand, out)
    e1.record()
    sync()
    launches = int(lib.bo_launch_count(1))
    pr = prof.stop()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())

    # ---- assertion 1: sharded top-k == single-shot top-k of the gathered scores, bit for bit
    kk = k + 16
    if world > 1:
        per = bd.shard_range(total, world, 0)[1]
        pad = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
        pad[: hi - lo] = out["acq"]
        full = bd.all_gather_cat(pad)[:total].contiguous()  # ranks own consecutive blocks of `per` candidates
        lv, li = gp.topk(out["acq"], kk, lo)
        gv, gi = bd.gather_topk(lv, li)
        mv, mi = gp.topk_merge(gv, gi, kk)
    else:
        full = out["acq"]
        parts = [gp.topk(full[a:b], kk, a) for a, b in (bd.shard_range(total, 4, r) for r in range(4))]
        mv, mi = gp.topk_merge(torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), kk)
    sv, si = gp.topk(full, kk, 0)
    same = bool(torch.equal(si, mi) and torch.equal(sv, mv))
    del full
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    # ---- assertion 2: oracle spot check of 256 candidates of every rank (rank 0 runs the CPU oracle once per config)
    chk = oracle_spot_check(out, cand, x, y, mu0, var0, ls, n, m, 256, cfg["cond"], fit_cache.setdefault(tag, {}),
                            world, rank)
    flops = float(total) * m * n * n
    ceiling = peaks["dgemm_tflops"] * 1e12 * world / (m * float(n) * n) if peaks["dgemm_tflops"] else None
    blk = {"variance_engine": engine, "value": total / secs, "unit": UNIT, "ms_per_step": 1e3 * secs,
           "scaling": "strong", "n_gpus": world, "candidates_total": total, "candidates_this_rank": hi - lo,
           "algorithmic_tflops_per_gpu": flops / secs / 1e12 / world, "gpu_launches_rank0": launches,
           "fp64_ceiling_candidates_per_s": ceiling,
           "frac_of_fp64_ceiling": (total / secs / ceiling) if ceiling else None,
           "rooflines_rank0": kernel_rooflines(pr, engine, secs, peaks),
           "sharded_topk_equals_gathered": bool(flag.item()),
           "oracle_spot_check": chk,
           "batch_idx": idx.cpu().tolist(), "batch_val": vals.cpu().tolist()}
    if engine == "int8":
        blk["guard"] = {"sampled_max_dvar_over_var0": gp.last_guard_worst, "tolerance": gp.last_guard_tolerance}
    if cfg.get("pareto"):
        fr = front.to(torch.int64)
        if world > 1:
            dist.all_reduce(fr)
        blk["pareto_front_of_ucb_vectors"] = int(fr.item())
    del cand, out, gp
    torch.cuda.empty_cache()
    return blk


def run_cfg5(dev, world, rank, peaks, lib):
    """cfg5: the 256-setting log-marginal-likelihood sweep at N=4096; settings are independent -> sharded."""
    import torch
    import torch.distributed as dist

    from bayesopt_smart_b200 import distributed as bd
    from bayesopt_smart_b200 import numba_kernels as nk
    from bayesopt_smart_b200.workloads import cfg5_settings, make_training_set

    cfg = CONFIGS["cfg5"]
    n, d, m = cfg["n"], cfg["d"], cfg["m"]
    x, y, mu0, _ = make_training_set(cfg["fn"], n, d, seed=0)
    ls, jit = cfg5_settings()
    lo, hi = bd.shard_range(len(ls), world, rank)
    ls2 = np.stack([ls, ls], axis=1)
    nk.mll_batched(x, y, mu0, ls2[lo:lo + 2], jit[lo:lo + 2], n)  # warm-up (workspace, kernels)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vals = nk.mll_batched(x, y, mu0, ls2[lo:hi], jit[lo:hi], n) if hi > lo else np.zeros(0)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    per = bd.shard_range(len(ls), world, 0)[1]
    pad = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
    pad[: hi - lo] = torch.from_numpy(vals).to(dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allv = bd.all_gather_cat(pad)[: len(ls)].cpu().numpy()
    else:
        allv = pad[: len(ls)].cpu().numpy()
    secs = float(t.item())
    flops = len(ls) * m * (n ** 3) / 3.0
    tf = flops / secs / 1e12 / world
    best = int(np.nanargmax(allv))
    return {"workload": cfg["name"], "settings": len(ls), "n_train": n, "seconds": secs, "n_gpus": world,
            "settings_per_s": len(ls) / secs, "all_finite": bool(np.isfinite(allv).all()),
            "best_setting": {"length_scale": float(ls[best]), "jitter": float(jit[best]), "mll": float(allv[best])},
            "roofline": {"kernel": "blocked Cholesky (gemm64_kernel trailing updates) + mll_solve_kernel",
                         "bound": "tensor", "achieved": tf, "peak": peaks["dgemm_tflops"], "unit": "TFLOP/s per GPU",
                         "frac": tf / peaks["dgemm_tflops"] if peaks["dgemm_tflops"] else None,
                         "algorithmic": "S * m * N^3 / 3 flop (potrf), Gram and the solves not counted"}}


def run_hbm_passes(dev, peaks, lib):
    """The memory-bound passes on their own (rank 0): stand-alone a6..a8 score pass and the top-k scan."""
    import torch

    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200.engine import DeviceGP, _ptr, _stream

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def best_of(fn, reps=5):
        fn()
        best = 1e30
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    out = {}
    m, big = 2, 16_000_000
    mu = torch.randn(m, big, dtype=torch.float64, device=dev)
    var = torch.rand(m, big, dtype=torch.float64, device=dev)
    smu, svar, ucb = torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(mu)
    acq = torch.empty(big, dtype=torch.float64, device=dev)
    _, pm = _lib.host_doubles(np.zeros(m), m)
    _, pv = _lib.host_doubles(np.full(m, 2.0), m)
    _, pb = _lib.host_doubles(np.full(m, 2.0), m)
    t = best_of(lambda: _lib.check(lib.bo_acquisition_f64(_ptr(smu), _ptr(svar), _ptr(ucb), _ptr(acq), _ptr(mu),
                                                          _ptr(var), big, big, m, pm, pv, pb, _stream())))
    nbytes = big * 8 * (2 * m + 3 * m + 1)
    out["score_pass_m2_16M"] = {"kernel": "acquisition_kernel", "bound": "hbm", "bytes": nbytes, "seconds": t,
                                "achieved": nbytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": nbytes / t / 1e9 / peaks["hbm_gbs"], "candidates_per_s": big / t,
                                "algorithmic": "88 B per candidate (read 2m, write 3m+1 doubles)"}
    # the same pass with the EXACT 2-objective HVI fused in (opt-in acquisition): standardise + UCB + HVI against a
    # device-prepared 300-point front, all five arrays written -> the same 88 B per candidate
    from bayesopt_smart_b200.engine import HviFront

    rng = np.random.default_rng(0)
    tt = np.sort(rng.random(300))
    front = HviFront(np.stack([tt, 1.0 - tt ** 2], axis=1) * 3.0 - 1.0, np.array([-1.5, -1.5]), dev)
    t = best_of(lambda: _lib.check(lib.bo_acquisition_hvi_f64(_ptr(smu), _ptr(svar), _ptr(ucb), _ptr(acq), _ptr(mu),
                                                              _ptr(var), big, big, m, pm, pv, pb, _ptr(front.prepared),
                                                              _ptr(front.count), front.n_points, front.ref_ptr(),
                                                              _stream())))
    out["fused_ucb_exact_hvi_m2_16M"] = {"kernel": "acquisition_hvi_kernel<2>", "bound": "hbm", "bytes": nbytes,
                                         "seconds": t, "achieved": nbytes / t / 1e9, "peak": peaks["hbm_gbs"],
                                         "unit": "GB/s", "frac": nbytes / t / 1e9 / peaks["hbm_gbs"],
                                         "candidates_per_s": big / t, "front_points": int(front.count.item()),
                                         "algorithmic": "88 B per candidate; HVI = two binary searches over the "
                                                        "staircase + O(1) (prefix areas), front tables in L1"}
    gp = DeviceGP(dev)
    acq.copy_(torch.randn(big, dtype=torch.float64, device=dev))
    t = best_of(lambda: gp.topk(acq, 19))
    out["topk_k19_16M"] = {"kernel": "topk_slice_kernel + topk_filter_kernel", "bound": "hbm", "bytes": big * 8,
                           "seconds": t, "achieved": big * 8 / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": big * 8 / t / 1e9 / peaks["hbm_gbs"],
                           "note": "8 B per score; the sample pass reads 1/stride of them again"}
    # a10 at BASELINE config 4's size: non-dominated filter of 8 M three-objective vectors, all inside the library
    from bayesopt_smart_b200.pareto import pareto_mask_device

    g = torch.Generator(device=dev).manual_seed(3)
    yv = torch.randn(8_000_000, 3, dtype=torch.float64, device=dev, generator=g)
    t = best_of(lambda: pareto_mask_device(yv), reps=3)
    out["pareto_filter_8M_x_3"] = {"kernel": "pareto_kernel (sample fronts, thinning, compaction, final test)",
                                   "seconds": t, "points_per_s": 8_000_000 / t,
                                   "front_points": int(pareto_mask_device(yv).sum().item()),
                                   "bytes_first_pass": 8_000_000 * 24, "first_pass_floor_s": 8_000_000 * 24 / (peaks["hbm_gbs"] * 1e9)}
    out["peak_source"] = peaks["hbm_source"]
    return out


def run_incremental_fit(dev):
    """time saved per iteration when the resident factor is extended by a batch of 3 new points
    (bo_gp_append_f64) instead of rebuilt (bo_gp_fit_f64) -- rank 0, CUDA events, best of 3."""
    import torch

    from bayesopt_smart_b200.engine import DeviceGP
    from bayesopt_smart_b200.workloads import make_training_set

    out = {}
    for n in (1024, 4096):
        x, y, mu0, var0 = make_training_set("zdt1", n, 6, seed=0)
        ls = np.full(2, 0.3)
        gp = DeviceGP(dev)
        full, app = [], []
        for _ in range(4):
            gp.fit(x, y, mu0, var0, ls, n - 3, incremental=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gp.fit(x, y, mu0, var0, ls, n)
            e1.record()
            torch.cuda.synchronize()
            assert gp.last_fit == "append"
            app.append(e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gp.fit(x, y, mu0, var0, ls, n, incremental=False)
            e1.record()
            torch.cuda.synchronize()
            full.append(e0.elapsed_time(e1))
        out[f"n{n}"] = {"full_refit_ms": min(full[1:]), "append_3_rows_ms": min(app[1:]),
                        "speedup": min(full[1:]) / min(app[1:])}
    out["what"] = "DeviceGP.fit with an unchanged prefix and bit-identical hyper-parameters extends L, W = L^-1 " \
                  "and alpha by the new rows (O(b N^2)); the reference rebuilds everything (O(N^3))"
    return out


def run_cfg1_loop():
    """BASELINE configs[0] through the drop-in class, Powell hyper-parameter fit included (a
    full-iteration number next to the hot-path figure)."""
    import bayesopt_smart_b200 as bo
    from bayesopt_smart_b200.workloads import toy_function

    mon = bo.PerformanceMonitor()
    np.random.seed(42)
    t0 = time.perf_counter()
    opt = bo.BayesianOptimization(function=toy_function, bounds=[(0, 300), (0, 300)], n_objectives=2, n_iterations=20,
                                  initial_samples=10, callbacks=[mon])
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        opt.optimize()
        front = opt.pareto_analysis()
    wall = time.perf_counter() - t0
    steady = {key: float(np.mean(v[1:])) for key, v in mon.timings.items()}
    return {"workload": "cfg1_demo_2d_toy_function_init10_iter20_batch3 (M = 90 000 int64 grid)", "wall_s": wall,
            "iterations": len(mon.timings["total"]), "steady_state_avg_s": steady,
            "hot_path_candidates_per_s_incl_powell": 90000 / steady["total"] if steady["total"] > 0 else None,
            "pareto_front": np.asarray(front).tolist(),
            "reference_measured_at_survey": {"avg_iteration_s": 2.38, "wall_s_incl_jit": 47.7, "cores": 8}}


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    t_start = time.perf_counter()
    import torch
    import torch.distributed as dist

    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200 import distributed as bd
    from bayesopt_smart_b200.engine import DeviceGP, PinnedMirror, hot_path_iteration, to_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    w = WORKLOAD
    n, m, k = w["n"], w["m"], w["batch"]
    x, y, mu0, var0, cand, ls, betas = make_workload(rank)
    n_cand = cand.shape[0]
    index_base = rank * n_cand

    gp = DeviceGP(dev, variance_engine=args.engine)
    x_dev, y_dev, cand_dev = to_device(x, device=dev), to_device(y, device=dev), to_device(cand, device=dev)
    out = {"acq": torch.empty(n_cand, dtype=torch.float64, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def exchange(vals, idx):
        """All-gather each rank's top-k pairs and merge with the same comparator (identical on every rank)."""
        if world == 1:
            return vals, idx
        gv, gi = bd.gather_topk(vals, idx)
        return bd.merge_topk(gp, gv, gi, vals.numel())

    def step_resident(g=None):
        """a1..a9 with every input already resident in HBM."""
        g = g or gp
        g.fit(x_dev, y_dev, mu0, var0, ls, n)
        g.score(cand_dev, betas, want=("acq",), out=out)
        vals, idx = g.topk(out["acq"], k + 16, index_base)
        flags = g.match_rows(idx, cand_dev, x_dev, index_base)
        vals = torch.where(flags.bool(), torch.full_like(vals, float("-inf")), vals)
        return exchange(vals, idx)

    # pinned host copies for the end-to-end variant (the call a user of the reference API makes)
    mirror = PinnedMirror()
    x_h, y_h, cand_h = (torch.from_numpy(a).pin_memory() for a in (x, y, cand))

    def step_e2e(g=None):
        """Same step through the public host-buffer API: H2D of x, y, candidates; D2H of mu, var, acq, batch."""
        res = hot_path_iteration(g or gp, x_h, y_h, cand_h, mu0, var0, ls, betas, n, k, mirror=mirror,
                                 index_base=index_base)
        if world > 1:
            exchange(res["top_vals_dev"], res["top_idx_dev"])
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, steps):
        """Per-step CUDA events on the launching stream; L2 flushed (untimed) between steps."""
        evs = []
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - wall0
        secs = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
        t = torch.tensor([secs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    # FP64 tensor peak: cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json has no FP64 entry)
    def dgemm_peak():
        nn = 8192
        a = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        b = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        best = 1e30
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1) * 1e-3)
        return 2.0 * nn**3 / best / 1e12

    hbm_gbs, hbm_src = measured_hbm_peak()
    peaks = {"dgemm_tflops": 0.0, "i8_tops": 0.0, "i8_burst_tops": 0.0, "hbm_gbs": hbm_gbs, "hbm_source": hbm_src}
    if not args.profile_mode:
        # every rank measures its own GPU (the extra blocks use them); rank 0's go into the headline roofline
        peaks["dgemm_tflops"] = dgemm_peak()
        pk, pkb = ctypes.c_double(), ctypes.c_double()
        _lib.check(lib.bo_i8_peak_tops(ctypes.byref(pkb), 0.03, None))
        _lib.check(lib.bo_i8_peak_tops(ctypes.byref(pk), 0.4, None))  # sustained: the figure kernels in long steps are held to
        peaks["i8_tops"], peaks["i8_burst_tops"] = pk.value, pkb.value
    peak_tflops, peak_i8_tops = peaks["dgemm_tflops"], peaks["i8_tops"]

    for _ in range(max(args.warmup, 3)):
        step_resident()
    warm = max(args.warmup, 3)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.bo_launch_count(1)
    prof = Profile(lib)
    prof.start()
    secs, wall = timed_steps(step_resident, args.steps)
    pr = prof.stop()
    ms_c, nl_c, fl_c = pr["contraction"]
    launches = int(lib.bo_launch_count(1))
    clock_info = clocks.stop() if rank == 0 else {}

    # end-to-end (host buffers in, host results out) -- same metric, copies inside the timed region
    if args.profile_mode:
        secs_e2e = float("nan")
    else:
        for _ in range(2):
            step_e2e()
        secs_e2e, _ = timed_steps(step_e2e, args.steps)
    h2d = x_h.numel() * 8 + y_h.numel() * 8 + cand_h.numel() * 8
    d2h = (2 * m + 1) * n_cand * 8 + (k + 16) * 16

    # ---- the same step with the INT8 tensor-core variance engine (reported next to the headline, not as it)
    i8 = None
    if not args.profile_mode and not args.no_int8 and args.engine == "dmma":
        acq_dmma = out["acq"].clone()
        top_dmma = step_resident()[1].clone()
        gp8 = DeviceGP(dev, variance_engine="int8")
        for _ in range(3):
            step_resident(gp8)
        lib.bo_launch_count(1)
        prof.start()
        secs8, _ = timed_steps(lambda: step_resident(gp8), args.steps)
        pr8 = prof.stop()
        launches8 = int(lib.bo_launch_count(1))
        top8 = step_resident(gp8)[1]
        dacq = float((out["acq"] - acq_dmma).abs().max().item())
        same_topk = bool(torch.equal(top8[:k], top_dmma[:k]))
        for _ in range(2):
            step_e2e(gp8)
        secs8_e2e, _ = timed_steps(lambda: step_e2e(gp8), args.steps)
        i8 = dict(secs=secs8, secs_e2e=secs8_e2e, prof=pr8, gpu_launches=launches8, dacq=dacq, same_topk=same_topk,
                  guard={"sampled_max_dvar_over_var0": gp8.last_guard_worst, "tolerance": gp8.last_guard_tolerance,
                         "stride": gp8.int8_guard_stride,
                         "what": "every INT8 scoring pass (inside the timed steps too) scores one candidate per "
                                 "stride with the FP64 engine as well and raises if they differ by more than the "
                                 "tolerance max(1e-9, 10 eps cond_upper)"})
        del gp8, acq_dmma

    # ---- the other BASELINE configs as whole jobs (every rank takes part; rank 0 reports)
    extras = None
    if not args.profile_mode and not args.no_extras:
        del cand_dev, out, flush
        torch.cuda.empty_cache()
        extras, fit_cache = {}, {}
        jobs = [("north_star", "hl"), ("cfg4", "cfg4")] + ([("cfg3", "cfg3")] if world >= 2 else [])
        for label, tag in jobs:
            extras[label] = {"workload": CONFIGS[tag]["name"], "engines": {}}
            for engine in ("dmma", "int8"):
                t_blk = time.perf_counter()
                try:
                    extras[label]["engines"][engine] = run_sharded_job(tag, engine, dev, world, rank, peaks, lib,
                                                                       fit_cache)
                except Exception as exc:  # noqa: BLE001  (reported, never hidden; the headline line still prints)
                    extras[label]["engines"][engine] = {"error": f"{type(exc).__name__}: {exc}"}
                extras[label]["engines"][engine]["block_wall_s"] = time.perf_counter() - t_blk  # incl. set-up + checks
        try:
            extras["cfg5"] = run_cfg5(dev, world, rank, peaks, lib)
        except Exception as exc:  # noqa: BLE001
            extras["cfg5"] = {"error": f"{type(exc).__name__}: {exc}"}
        if rank == 0:
            for label, fn in (("hbm_passes", lambda: run_hbm_passes(dev, peaks, lib)),
                              ("incremental_fit", lambda: run_incremental_fit(dev)), ("cfg1_loop", run_cfg1_loop)):
                if world > 1 and label == "cfg1_loop":
                    continue  # BayesianOptimization.optimize() would shard over the process group
                try:
                    extras[label] = fn()
                except Exception as exc:  # noqa: BLE001
                    extras[label] = {"error": f"{type(exc).__name__}: {exc}"}
        if world > 1:
            dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def load_traffic(name):
        """DRAM bytes per launch from the committed `ncu --set full` capture of the same kernel / workload"""
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f)
            if tr["n_train"] == n and tr["objectives"] == m:
                return tr
        except (OSError, KeyError, ValueError):
            pass
        return None

    def roofline_of(engine, pr_, secs_, peak_dgemm, peak_i8, peak_i8_burst=None):
        ms_, nl_, fl_ = pr_["contraction"]
        tf = fl_ / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
        tr = load_traffic("trmm_traffic.json" if engine == "dmma" else "oz_traffic.json")
        cands_per_launch = (fl_ / nl_) / (m * float(n) * n) if nl_ else 0.0
        common = {"launches": nl_, "avg_launch_ms": ms_ / max(1, nl_), "share_of_step": ms_ * 1e-3 / secs_,
                  "candidates_per_launch": cands_per_launch,
                  # flat: DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum), null if no capture
                  "traffic": tr["dram_bytes_per_launch"] if tr else None,
                  "traffic_source": ("profiles/" + ("trmm_traffic.json" if engine == "dmma" else "oz_traffic.json"))
                  if tr else None,
                  "algorithmic_bytes_per_launch": cands_per_launch * (8.0 * w["d"] + 8.0 * (5 * m + 1)),
                  "other_kernels": {kname: v for kname, v in kernel_rooflines(pr_, engine, secs_, peaks).items()
                                    if "sumsq" not in kname}}
        if engine == "dmma":
            return {"kernel": "trmm_sumsq_kernel", "bound": "tensor", "achieved": tf, "peak": peak_dgemm,
                    "unit": "TFLOP/s", "frac": tf / peak_dgemm if peak_dgemm else None,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 is absent from "
                                   "MEASURED_PEAKS.json; nominal B200 FP64 tensor 37-40 TFLOP/s)",
                    "algorithmic": "m*N^2 flop per candidate = 2.097e6; per launch x candidates in the chunk", **common}
        return {"kernel": "oz_sumsq_kernel", "bound": "tensor", "achieved": 21.0 * tf, "peak": peak_i8,
                "peak_burst": peak_i8_burst, "unit": "TOP/s (int8 multiply + add)",
                "frac": 21.0 * tf / peak_i8 if peak_i8 else None, "fp64_equivalent_tflops": tf,
                "peak_source": "bo_i8_peak_tops: the kernel's own 21-MMA batch (kind::i8 128x64x32, A from TMEM) on "
                               "resident operands, one CTA per SM, measured in this run: `peak` sustained for 0.4 s "
                               "(the contraction is timed inside long steps), `peak_burst` for 30 ms "
                               "(MEASURED_PEAKS.json has no int8 entry; nominal B200 dense int8 4.5 POP/s)",
                "algorithmic": "21 digit-pair products x m*N^2 multiply-adds per candidate", **common}

    total_cands = n_cand * world * args.steps
    # bounded CPU sample of the same workload: the reference's own Numba path (oracle/_ref), ~15-30 s
    cpu_block = None
    if not args.profile_mode and world == 1:  # the CPU baseline is reported at N = 1 only
        try:
            ref = CpuReference()
            cpu_block = ref.baseline_block(ref.measure(sample_cands=2 * CPU_CHUNK, steps=3, warmup=1))
        except Exception as exc:  # noqa: BLE001
            cpu_block = {"error": f"{type(exc).__name__}: {exc}"}
    cfg = headline_config(world)
    line = {
        "metric": METRIC, "value": total_cands / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "variance_engine": args.engine,
        "config": cfg, "wall_s_incl_flush": wall,
        "clocks": clock_info,
        "e2e": {"value": total_cands / secs_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * secs_e2e / args.steps,
                "api": "engine.hot_path_iteration (pinned host x, y, input_space in; mu, var, acq, batch out)"},
        "gpu_launches": launches,
        "roofline": roofline_of(args.engine, pr, secs, peak_tflops, peak_i8_tops, peaks["i8_burst_tops"]),
        "int8_engine": (None if i8 is None else {
            "what": "same step with variance_engine='int8': |W k*|^2 by error-free splitting into 6 balanced "
                    "base-256 digit planes, 21 digit-pair products on tcgen05.mma.kind::i8 (exact int32 in TMEM), "
                    "int64/FP64 recombination; everything else identical",
            "value": total_cands / i8["secs"], "unit": UNIT, "ms_per_step": 1e3 * i8["secs"] / args.steps,
            "e2e": {"value": total_cands / i8["secs_e2e"], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * i8["secs_e2e"] / args.steps},
            "gpu_launches": i8["gpu_launches"],
            "max_abs_acq_difference_vs_dmma": i8["dacq"], "same_top_batch_as_dmma": i8["same_topk"],
            "guard": i8["guard"],
            "roofline": roofline_of("int8", i8["prof"], i8["secs"], peak_tflops, peak_i8_tops, peaks["i8_burst_tops"])}),
        "baseline_configs": extras,
        "cpu_baseline": cpu_block,
        "bench_wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-int8", action="store_true", help="skip the second measurement with the INT8 engine")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip the other BASELINE configs (north star shape, cfg3, cfg4, cfg5, ...)")
    ap.add_argument("--engine", default="dmma", choices=["dmma", "int8"],
                    help="variance engine of the headline measurement (default: the FP64 DMMA engine; with int8 the "
                         "whole line, roofline included, describes the INT8 engine)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="short run for ncu: skips the peak probes, the end-to-end leg, the extras and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
