"""Time the training-side factorisation (bo_gp_fit_f64) and the cfg5 batched MLL sweep.
Usage: python tools/fit_time.py [sweep_settings]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200 import numba_kernels as nk  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)

for n, d in ((1024, 6), (2048, 8), (4096, 6)):
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=0)
    gp = DeviceGP()
    ls = np.full(2, 0.3)
    for _ in range(2):
        gp.fit(x, y, mu0, var0, ls, n)
    xd, yd = gp.x, torch.from_numpy(y).cuda()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gp.fit(xd, yd, mu0, var0, ls, n)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(json.dumps(dict(kind="fit", n=n, d=d, fit_ms=float(np.median(ts)))), flush=True)

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, d, m = 4096, 6, 2
x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=0)
ls = np.repeat(np.logspace(-1, 0.5, 16), 16)[:S]
jit = np.tile(np.logspace(-8, -2, 16), 16)[:S]
vals = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1), jit, n)
t0 = time.perf_counter()
vals = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1), jit, n)
t = time.perf_counter() - t0
print(json.dumps(dict(kind="mll_sweep", settings=S, n=n, seconds=t, potrf_tflops=S * m * n**3 / 3.0 / t / 1e12,
                      nan=int(np.isnan(vals).sum()), value_setting0=float(vals[0]))))
