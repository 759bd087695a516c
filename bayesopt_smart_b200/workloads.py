"""Synthetic workloads of the BASELINE configs: ZDT1 / ZDT2 / DTLZ2 training sets.

Pure NumPy, no arithmetic of the hot path: these only DEFINE the inputs that ``bench.py`` and the tools under
``tools/`` feed to the CUDA path (the reference ships its own toy objectives in ``bayesopt/benchmark_functions.py``;
the multi-objective test functions named by BASELINE.json are the standard ones).  The reference maximises, so the
objectives are negated; ``prior_mean`` / ``prior_variance`` are the sample mean / variance, as the reference's
``compute_prior_mean`` / ``compute_prior_variance`` would set them.
"""
from __future__ import annotations

import numpy as np


def zdt1(x: np.ndarray) -> np.ndarray:
    g = 1.0 + 9.0 * np.mean(x[:, 1:], axis=1)
    f1 = x[:, 0]
    return np.stack([f1, g * (1.0 - np.sqrt(f1 / g))], axis=1)


def zdt2(x: np.ndarray) -> np.ndarray:
    g = 1.0 + 9.0 * np.mean(x[:, 1:], axis=1)
    f1 = x[:, 0]
    return np.stack([f1, g * (1.0 - (f1 / g) ** 2)], axis=1)


def dtlz2(x: np.ndarray, n_obj: int = 3) -> np.ndarray:
    g = np.sum((x[:, n_obj - 1:] - 0.5) ** 2, axis=1)
    out = []
    for i in range(n_obj):
        f = 1.0 + g
        for j in range(n_obj - 1 - i):
            f = f * np.cos(0.5 * np.pi * x[:, j])
        if i > 0:
            f = f * np.sin(0.5 * np.pi * x[:, n_obj - 1 - i])
        out.append(f)
    return np.stack(out, axis=1)


def make_training_set(name: str, n: int, d: int, seed: int = 0):
    """X ~ U[0,1]^{n x d}; y = -objective (the reference maximises); mu0 = mean, var0 = var."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    fn = {"zdt1": zdt1, "zdt2": zdt2, "dtlz2": dtlz2}[name]
    y = -fn(x)
    return x, y, y.mean(axis=0), y.var(axis=0)
