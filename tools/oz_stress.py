"""Randomised stress test of the INT8 engine: many shapes, each scored twice (bit-identical?) and compared with
the FP64 DMMA engine (tolerance max(1e-9, 10 eps cond), the same as in the parity tests).  Usage: python tools/oz_stress.py [cases] [seed]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst, nondet, fails = 0.0, 0, 0
for c in range(cases):
    n = int(rng.choice([1, 7, 63, 64, 65, 127, 128, 129, 200, 256, 300, 511, 512, 640, 1000, 1024, 1500, 2048]))
    d = int(rng.integers(1, 17))
    m = int(rng.integers(1, 5))
    n_cand = int(rng.choice([1, 63, 64, 65, 500, 4096, 4097, 75776, 75777, 120000, 200000]))
    ls = float(rng.uniform(0.25, 1.0) * np.sqrt(d) / 2)
    x = rng.random((n, d))
    w = rng.normal(size=(d, m))
    y = np.sin(2.0 * x @ w) + 0.05 * rng.normal(size=(n, m))
    mu0, var0 = y.mean(0), np.maximum(y.var(0), 0.05)
    cand = to_device(rng.random((n_cand, d)))
    betas = rng.uniform(0.5, 2.5, size=m)
    g8, gd = DeviceGP(variance_engine="int8"), DeviceGP(variance_engine="dmma")
    g8.fit(x, y, mu0, var0, np.full(m, ls), n)
    gd.fit(x, y, mu0, var0, np.full(m, ls), n)
    a = g8.score(cand, betas, want=("mu", "var", "acq"))
    a = {k: v.clone() for k, v in a.items()}
    b = g8.score(cand, betas, want=("mu", "var", "acq"))
    r = gd.score(cand, betas, want=("mu", "var", "acq"))
    torch.cuda.synchronize()
    same = all(torch.equal(a[k], b[k]) for k in a)
    dv = float(((a["var"] - r["var"]).abs().amax(dim=1).cpu().numpy() / var0).max())
    dm = float(((a["mu"] - r["mu"]).abs().amax(dim=1).cpu().numpy() / np.sqrt(var0)).max())
    tau = 1e-9
    if dv >= tau or dm >= tau:  # the tolerance widens with the conditioning of K + 1e-6 I
        sq = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
        cond = max(np.linalg.cond(var0[o] * np.exp(-0.5 * sq / ls ** 2) + 1e-6 * np.eye(n)) for o in range(m))
        tau = max(1e-9, 10 * np.finfo(np.float64).eps * cond)
    ok = same and dv < tau and dm < tau and bool(torch.isfinite(a["acq"]).all())
    worst = max(worst, dv)
    nondet += not same
    fails += not ok
    if not ok:
        print(json.dumps(dict(case=c, n=n, d=d, m=m, n_cand=n_cand, ls=ls, same=same, dvar=dv, dmu=dm)), flush=True)
print(json.dumps(dict(cases=cases, failures=fails, nondeterministic=nondet, worst_dvar_over_var0=worst)))
sys.exit(1 if fails else 0)
