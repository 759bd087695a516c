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


def toy_function(x):
    """The 2-objective toy problem of BASELINE config 1 (reference examples/benchmark_functions.py:33-50):
    maxima at x0 = 150 (f = 100) and x1 = 150 (g = 20)."""
    return np.array([-((x[0] - 150) ** 2) + 100, -((x[1] - 150) ** 2) + 20])


# The BASELINE.json configs as synthetic workloads (shapes, hyper-parameters, seeds).
# `cond` is the condition number of K + 1e-6 I measured at survey time for these hyper-parameters; the parity
# tolerance in standardised units is tau = max(1e-9, 10 * eps * cond)  (SURVEY 8(c)).
CONFIGS = {
    "cfg2": dict(name="cfg2_zdt1_d6_n1024_m2_grid1M", fn="zdt1", n=1024, d=6, m=2, ls=0.3, beta=2.0, batch=3,
                 grid_levels=10, total=1_000_000, cond=1.3e4),
    "cfg3": dict(name="cfg3_zdt2_d10_n4096_m2_rand16M", fn="zdt2", n=4096, d=10, m=2, ls=0.5, beta=2.0, batch=3,
                 total=16_000_000, cond=1.1e5),
    "cfg4": dict(name="cfg4_dtlz2_d8_n2048_m3_rand8M_pareto", fn="dtlz2", n=2048, d=8, m=3, ls=0.5, beta=2.0,
                 batch=3, total=8_000_000, pareto=True, cond=6.6e5),
    "cfg5": dict(name="cfg5_mll_sweep_256_settings_n4096_d6", fn="zdt1", n=4096, d=6, m=2, settings=256),
    # the north star's headline shape: N = 4096, d = 6, 2 objectives, >= 16 M candidates
    "hl": dict(name="northstar_zdt1_d6_n4096_m2_rand16M", fn="zdt1", n=4096, d=6, m=2, ls=0.3, beta=2.0, batch=3,
               total=16_000_000, cond=2.4e6),
}

CANDIDATE_CHUNK = 250_000  # candidates per generator chunk; the global chunk index seeds the generator


def shard_candidates(lo: int, hi: int, d: int, device):
    """Rows [lo, hi) of the synthetic candidate set U[0,1]^{M x d}, generated on the device.  Chunk c of 250 000
    rows comes from a generator seeded with c, so any rank count (and any shard boundary) sees the same global
    candidate set -- the property the sharded-vs-gathered checks rely on."""
    import torch

    parts = []
    for c in range(lo // CANDIDATE_CHUNK, (hi + CANDIDATE_CHUNK - 1) // CANDIDATE_CHUNK):
        g = torch.Generator(device=device).manual_seed(1234 + c)
        block = torch.rand(CANDIDATE_CHUNK, d, dtype=torch.float64, device=device, generator=g)
        a, b = max(lo, c * CANDIDATE_CHUNK) - c * CANDIDATE_CHUNK, min(hi, (c + 1) * CANDIDATE_CHUNK) - c * CANDIDATE_CHUNK
        parts.append(block[a:b])
    if not parts:
        return torch.empty((0, d), dtype=torch.float64, device=device)
    return torch.cat(parts).contiguous()


def cfg5_settings():
    """16 length scales in logspace(-1, 0.5) x 16 jitters ("noise") in logspace(-8, -2) = 256 settings."""
    ls = np.repeat(np.logspace(-1, 0.5, 16), 16)
    jit = np.tile(np.logspace(-8, -2, 16), 16)
    return ls, jit
