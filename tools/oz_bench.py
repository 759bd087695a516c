"""Device timing of the two variance engines on one config (CUDA events, warm-up, L2 flush between runs).
Usage: python tools/oz_bench.py [n] [n_cand] [d] [m] [reps]   -> JSON lines"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200 import _lib  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)


def timed(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    m = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    lib = _lib.load()
    name = {2: "zdt1", 3: "dtlz2"}.get(m, "zdt1")
    x, y, mu0, var0 = orc.make_training_set(name, n, d, seed=0)
    rng = np.random.default_rng(1)
    cand = to_device(rng.random((n_cand, d)))
    ls = np.full(m, 0.3 if d <= 6 else 0.5)
    betas = np.full(m, 2.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    outs = {}
    for eng in ("dmma", "int8"):
        gp = DeviceGP(variance_engine=eng)
        gp.fit(x, y, mu0, var0, ls, n)
        out = {k: torch.empty((n_cand,) if k == "acq" else (m, n_cand), dtype=torch.float64, device="cuda")
               for k in ("mu", "var", "acq")}
        for _ in range(2):
            gp.score(cand, betas, out=out)
        torch.cuda.synchronize()
        lib.bo_profile_enable(1)
        tot, lau, fl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.bo_profile_read(ctypes.byref(tot), ctypes.byref(lau), ctypes.byref(fl))
        med, mn = timed(lambda: gp.score(cand, betas, out=out), reps, flush)
        lib.bo_profile_read(ctypes.byref(tot), ctypes.byref(lau), ctypes.byref(fl))
        lib.bo_profile_enable(0)
        fit_med, _ = timed(lambda: gp.fit(x, y, mu0, var0, ls, n), 3, flush)
        outs[eng] = {k: v.clone() for k, v in out.items()}
        res[eng] = dict(engine=eng, n=n, d=d, m=m, n_cand=n_cand, score_ms_median=med, score_ms_min=mn,
                        cand_per_s=n_cand / (med * 1e-3), contraction_ms_per_pass=tot.value / reps,
                        contraction_launches_per_pass=lau.value / reps,
                        contraction_tflops_fp64_equiv=fl.value / (tot.value * 1e-3) / 1e12 if tot.value else None,
                        fit_ms=fit_med)
        print(json.dumps(res[eng]), flush=True)
    dv = (outs["int8"]["var"] - outs["dmma"]["var"]).abs().amax(dim=1).cpu().numpy() / var0
    da = (outs["int8"]["acq"] - outs["dmma"]["acq"]).abs().max().item()
    print(json.dumps(dict(kind="engines_agree", max_dvar_over_var0=dv.tolist(), max_dacq=da,
                          speedup_score=res["dmma"]["score_ms_median"] / res["int8"]["score_ms_median"])))


if __name__ == "__main__":
    main()
