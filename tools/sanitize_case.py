"""Small pass through every kernel of libbo_b200 for compute-sanitizer (memcheck / racecheck), sizes kept tiny."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesopt_smart_b200 as bo  # noqa: E402
from bayesopt_smart_b200 import acquisition as aq  # noqa: E402
from bayesopt_smart_b200 import numba_kernels as nk  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402

rng = np.random.default_rng(0)
n, d, m, M = 150, 5, 3, 700
x = rng.random((n, d))
y = np.sin(3 * x @ rng.normal(size=(d, m)))
mu0, var0 = y.mean(0), y.var(0)
ls, betas = np.full(m, 0.5), np.full(m, 2.0)
cand = rng.random((M, d))
gp = DeviceGP()
gp.fit(x, y, mu0, var0, ls, n)
out = gp.score(cand, betas, want=("mu", "var", "std_mu", "std_var", "ucb", "acq"))
vals, idx = gp.select(to_device(cand), out["acq"], to_device(x), 3)
k = np.zeros((m, n, n))
nk.update_k(k, x, 0, n, var0, ls)
kinv = nk.invert_k(n, k)
ks = np.zeros((m, n, M))
nk.update_k_star(ks, x, cand, 0, n, var0, ls)
mu = np.zeros((m, M))
var = np.zeros((m, M))
nk.update_mean(mu, ks, kinv, y, mu0, n)
nk.update_variance(var, ks, kinv, var0, n)
print("mll", nk.mll_batched(x, y, mu0, np.full((3, m), 0.5), [1e-8, 1e-6, 1e-4], n))
print("pareto", bo.is_pareto_efficient(rng.normal(size=(3000, 3))).sum())
front = y[bo.is_pareto_efficient(y)]
hv = np.zeros(M)
aq.update_exact_hypervolume_improvement(hv, out["ucb"].cpu().numpy(), (front - mu0) / np.sqrt(var0), -3 * np.ones(m))
torch.cuda.synchronize()
print("ok", idx, float(np.abs(mu - out["mu"].cpu().numpy()).max()), float(hv.max()))
