#!/usr/bin/env python
"""bench.py -- candidate scores/sec of the acquisition hot path (GP predict + UCB + sum-UCB "HVI"
+ top-k batch selection) on N B200s, next to the CPU restatement of the reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the reference's per-iteration hot path (bayesian_optimization.py:129-207:
update_k, invert_k, update_k_star, update_mean, update_variance, standardize_objectives, update_ucb,
update_hypervolume_improvement, select_next_batch) over one synthetic candidate set.

Workload (BASELINE.json configs[1], "cfg2"): ZDT1, d = 6, N = 1024 training points, 2 objectives,
10^6-point candidate grid linspace(0,1,10)^6 per GPU, length scale 0.3, beta 2, batch 3.
N > 1: weak scaling -- every rank scores its own 10^6 candidates (rank 0 the grid, rank r a
counter-seeded uniform shard); the factor is recomputed per rank (deterministic, no broadcast) and
the only exchange is an all-gather of each rank's top-k (value, global index) pairs over NCCL,
merged on device with the same total order.

Printed keys follow the driver contract; `roofline` describes trmm_sumsq_kernel (FP64 DMMA bound),
`cpu_baseline` the NumPy port of the reference (oracle/gp_oracle.py) on the host cores.

The headline numbers (`value`, `e2e`, `roofline`) are measured with the FP64 DMMA variance engine, the
one BASELINE.json's north star names.  The same step is then timed with the INT8 tensor-core engine
(error-free digit splitting on tcgen05.mma.kind::i8, `variance_engine="int8"`) and reported under
`int8_engine`, with its own roofline (executed against the kind::i8 rate measured in the same run) and
the largest difference between the two engines' outputs on this workload.
"""
from __future__ import annotations

import argparse
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

WORKLOAD = dict(name="cfg2_zdt1_d6_n1024_m2_grid1M", fn="zdt1", n=1024, d=6, m=2, ls=0.3, beta=2.0, batch=3,
                grid_levels=10)
METRIC = "candidate scores/sec (GP predict+UCB+HVI)"
UNIT = "candidates/s"


# --------------------------------------------------------------------------------------- workload
def make_workload(rank: int = 0):
    from bayesopt_smart_b200.workloads import make_training_set  # input definition only (no hot-path arithmetic)

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
def cpu_reference_sample(sample_cands: int, chunk: int = 16384):
    """The NumPy port of the reference path (oracle.ref_hot_path pieces) on the host cores.

    Runs fit once (update_k + invert_k) and the per-candidate stages on `sample_cands` candidates of the
    workload in chunks (the reference materialises k_star (m, N, M): a single shot does not fit), then the
    selection on the scored sample.  Returns seconds for the sample."""
    from oracle import gp_oracle as orc

    x, y, mu0, var0, cand, ls, betas = make_workload(0)
    w = WORKLOAD
    n, m = w["n"], w["m"]
    cand = cand[:: max(1, cand.shape[0] // sample_cands)][:sample_cands]
    t0 = time.perf_counter()
    kmat = np.zeros((m, n, n))
    orc.ref_update_k(kmat, x, 0, n, var0, ls)
    kinv = orc.ref_invert_k(n, kmat)
    acq = np.zeros(cand.shape[0])
    for c0 in range(0, cand.shape[0], chunk):
        cc = cand[c0:c0 + chunk]
        ks = np.zeros((m, n, cc.shape[0]))
        orc.ref_update_k_star(ks, x, cc, 0, n, var0, ls)
        mu = np.zeros((m, cc.shape[0]))
        var = np.zeros_like(mu)
        orc.ref_update_mean(mu, ks, kinv, y, mu0, n)
        orc.ref_update_variance(var, ks, kinv, var0, n)
        smu, svar, ucb = np.zeros_like(mu), np.zeros_like(mu), np.zeros_like(mu)
        orc.ref_standardize_objectives(smu, svar, mu, var, mu0, var0)
        orc.ref_update_ucb(ucb, smu, svar, betas)
        orc.ref_update_hypervolume_improvement(acq[c0:c0 + chunk], ucb)
    orc.ref_select_next_batch(cand, acq, x[:n], w["batch"])
    return time.perf_counter() - t0, cand.shape[0]


def use_all_host_threads() -> None:
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:  # noqa: BLE001
        pass


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    sample = 32768
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_sample(4096)
    times = []
    for _ in range(args.steps):
        t, cnt = cpu_reference_sample(sample)
        times.append(t)
    total = float(np.sum(times))
    value = sample * args.steps / total
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "n_train": WORKLOAD["n"], "dims": WORKLOAD["d"],
                   "objectives": WORKLOAD["m"], "sample": f"{sample} of 10^6 candidates per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} grid candidates per step x {args.steps} steps, NumPy/OpenBLAS port "
                                   "of the reference functions (oracle/gp_oracle.py), chunks of 16384"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
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

    peak_tflops = dgemm_peak() if (rank == 0 and not args.profile_mode) else 0.0
    peak_i8_tops = 0.0
    if args.engine == "int8" and rank == 0 and not args.profile_mode:
        import ctypes as _ct

        pk = _ct.c_double()
        _lib.check(lib.bo_i8_peak_tops(_ct.byref(pk), 0.4, None))  # sustained (see int8_engine.roofline)
        peak_i8_tops = pk.value

    for _ in range(max(args.warmup, 3)):
        step_resident()
    warm = max(args.warmup, 3)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.bo_launch_count(1)
    lib.bo_profile_enable(1)
    import ctypes

    lib.bo_profile_read(None, None, None)
    secs, wall = timed_steps(step_resident, args.steps)
    ms = ctypes.c_double()
    nl = ctypes.c_longlong()
    fl = ctypes.c_double()
    lib.bo_profile_read(ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(fl))
    lib.bo_profile_enable(0)
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
        lib.bo_profile_enable(1)
        lib.bo_profile_read(None, None, None)
        secs8, _ = timed_steps(lambda: step_resident(gp8), args.steps)
        ms8, nl8, fl8 = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.bo_profile_read(ctypes.byref(ms8), ctypes.byref(nl8), ctypes.byref(fl8))
        lib.bo_profile_enable(0)
        launches8 = int(lib.bo_launch_count(1))
        top8 = step_resident(gp8)[1]
        dacq = float((out["acq"] - acq_dmma).abs().max().item())
        same_topk = bool(torch.equal(top8[:k], top_dmma[:k]))
        for _ in range(2):
            step_e2e(gp8)
        secs8_e2e, _ = timed_steps(lambda: step_e2e(gp8), args.steps)
        # the roofline denominator: the kernel's own MMA batch on resident operands, as a 30 ms burst and sustained
        # for 0.4 s (the contraction is timed inside long steps, so the sustained figure is the one it is held to)
        peak8b, peak8 = ctypes.c_double(), ctypes.c_double()
        _lib.check(lib.bo_i8_peak_tops(ctypes.byref(peak8b), 0.03, None))
        _lib.check(lib.bo_i8_peak_tops(ctypes.byref(peak8), 0.4, None))
        i8 = dict(secs=secs8, secs_e2e=secs8_e2e, ms=ms8.value, launches=int(nl8.value), flops=fl8.value,
                  gpu_launches=launches8, dacq=dacq, same_topk=same_topk, peak_tops=peak8.value,
                  peak_burst_tops=peak8b.value)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def load_traffic(name):
        """DRAM bytes per launch from the committed ncu capture of the same kernel / workload"""
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f)
            if tr["n_train"] == n and tr["objectives"] == m:
                return {"bytes_per_launch": tr["dram_bytes_per_launch"],
                        "candidates_per_launch": tr["candidates_per_launch"], "source": "profiles/" + name}
        except (OSError, KeyError, ValueError):
            pass
        return None

    traffic = load_traffic("trmm_traffic.json" if args.engine == "dmma" else "oz_traffic.json")
    total_cands = n_cand * world * args.steps
    achieved = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
    # bounded CPU sample of the same workload (reference port), ~10-20 s
    if args.profile_mode or world > 1:  # the CPU baseline is reported at N = 1 only
        t_cpu, cnt_cpu = float("nan"), 0
    else:
        use_all_host_threads()
        t_cpu, cnt_cpu = cpu_reference_sample(131072)
    cpu_val = cnt_cpu / t_cpu if cnt_cpu else None
    line = {
        "metric": METRIC, "value": total_cands / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "variance_engine": args.engine,
        "config": {"workload": w["name"], "n_train": n, "dims": w["d"], "objectives": m,
                   "candidates_per_gpu": n_cand, "batch_size": k, "length_scale": w["ls"], "beta": w["beta"],
                   "step": "update_k+invert_k (Cholesky/W) + K* + mean + variance + standardise + UCB + sum-UCB "
                           "+ top-k with evaluated-row exclusion" + (" + NCCL all-gather/merge" if world > 1 else ""),
                   "l2": "256 MiB flush write between timed steps (untimed); K* staging per chunk 0.6 GB > L2",
                   "timing": "per-step CUDA events on the launching stream, max over ranks",
                   "wall_s_incl_flush": wall},
        "clocks": clock_info,
        "e2e": {"value": total_cands / secs_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * secs_e2e / args.steps,
                "api": "engine.hot_path_iteration (pinned host x, y, input_space in; mu, var, acq, batch out)"},
        "gpu_launches": launches,
        "roofline": ({"kernel": "trmm_sumsq_kernel", "bound": "tensor", "achieved": achieved, "peak": peak_tflops,
                      "unit": "TFLOP/s", "frac": achieved / peak_tflops if peak_tflops else None, "traffic": traffic,
                      "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 is absent from "
                                     "MEASURED_PEAKS.json; nominal B200 FP64 tensor 37-40 TFLOP/s)",
                      "algorithmic": "m*N^2 flop per candidate = 2.097e6; per launch x candidates in the chunk",
                      "launches": int(nl.value), "avg_launch_ms": ms.value / max(1, nl.value),
                      "share_of_step": ms.value * 1e-3 / secs} if args.engine == "dmma" else
                     {"kernel": "oz_sumsq_kernel", "bound": "tensor", "achieved": 21.0 * achieved,
                      "peak": peak_i8_tops, "unit": "TOP/s (int8 multiply + add)",
                      "frac": 21.0 * achieved / peak_i8_tops if peak_i8_tops else None, "traffic": traffic,
                      "fp64_equivalent_tflops": achieved,
                      "peak_source": "bo_i8_peak_tops: the kernel's own 21-MMA batch on resident operands, one CTA "
                                     "per SM, ~30 ms, measured in this run",
                      "algorithmic": "21 digit-pair products x m*N^2 multiply-adds per candidate",
                      "launches": int(nl.value), "avg_launch_ms": ms.value / max(1, nl.value),
                      "share_of_step": ms.value * 1e-3 / secs}),
        "int8_engine": (None if i8 is None else {
            "what": "same step with variance_engine='int8': |W k*|^2 by error-free splitting into 6 balanced "
                    "base-256 digit planes, 21 digit-pair products on tcgen05.mma.kind::i8 (exact int32 in TMEM), "
                    "int64/FP64 recombination; everything else identical",
            "value": total_cands / i8["secs"], "unit": UNIT, "ms_per_step": 1e3 * i8["secs"] / args.steps,
            "e2e": {"value": total_cands / i8["secs_e2e"], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * i8["secs_e2e"] / args.steps},
            "gpu_launches": i8["gpu_launches"],
            "max_abs_acq_difference_vs_dmma": i8["dacq"], "same_top_batch_as_dmma": i8["same_topk"],
            "roofline": {"kernel": "oz_sumsq_kernel", "bound": "tensor",
                         "achieved": 21.0 * i8["flops"] / (i8["ms"] * 1e-3) / 1e12 if i8["ms"] > 0 else None,
                         "peak": i8["peak_tops"], "peak_burst": i8["peak_burst_tops"],
                         "unit": "TOP/s (int8 multiply + add)",
                         "frac": (21.0 * i8["flops"] / (i8["ms"] * 1e-3) / 1e12 / i8["peak_tops"])
                         if i8["ms"] > 0 and i8["peak_tops"] else None,
                         "fp64_equivalent_tflops": i8["flops"] / (i8["ms"] * 1e-3) / 1e12 if i8["ms"] > 0 else None,
                         "peak_source": "bo_i8_peak_tops: the kernel's own 21-MMA batch (kind::i8 128x64x32, A from "
                                        "TMEM) on resident operands, one CTA per SM, measured in this run: `peak` "
                                        "sustained for 0.4 s (the contraction is timed inside long steps), "
                                        "`peak_burst` for 30 ms (MEASURED_PEAKS.json has no int8 entry; nominal "
                                        "B200 dense int8 4.5 POP/s)",
                         "algorithmic": "21 digit-pair products x m*N^2 multiply-adds per candidate",
                         "traffic": load_traffic("oz_traffic.json"),
                         "launches": i8["launches"], "avg_launch_ms": i8["ms"] / max(1, i8["launches"]),
                         "share_of_step": i8["ms"] * 1e-3 / i8["secs"]}}),
        "cpu_baseline": ({"value": cpu_val, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                          "sample": f"{cnt_cpu} of 10^6 grid candidates, one pass, NumPy/OpenBLAS port of the "
                                    "reference functions (oracle/gp_oracle.py)"} if cnt_cpu else None),
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
    ap.add_argument("--engine", default="dmma", choices=["dmma", "int8"],
                    help="variance engine of the headline measurement (default: the FP64 DMMA engine; with int8 the "
                         "whole line, roofline included, describes the INT8 engine)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="short run for ncu: skips the DGEMM peak probe, the end-to-end leg and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
