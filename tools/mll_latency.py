"""Latency of one log-marginal-likelihood evaluation (what SciPy Powell calls hundreds of times per BO iteration)
as a function of the training-set size.  Usage: python tools/mll_latency.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200 import numba_kernels as nk  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)

for n in (64, 128, 129, 192, 256, 384, 512, 768, 1024, 2048):
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, 6, seed=0)
    ls = np.array([[0.3, 0.3]])
    jit = [1e-8]
    for _ in range(5):
        nk.mll_batched(x, y, mu0, ls, jit, n)
    torch.cuda.synchronize()
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        v = nk.mll_batched(x, y, mu0, ls, jit, n)
    t = (time.perf_counter() - t0) / reps
    print(json.dumps(dict(n=n, us_per_eval_incl_host=1e6 * t, value=float(v[0]))), flush=True)
