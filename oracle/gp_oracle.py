"""CPU oracle for the acquisition hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a plain NumPy restatement of the algorithm that
alebal123bal/BayesOpt_smart runs on its hot path (GP posterior over a candidate
set -> UCB -> sum-UCB "HVI" -> top-k batch -> Pareto filter).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the shipped package ``bayesopt_smart_b200`` never
does (it fails loudly when its CUDA library is missing).

Parity status
-------------
* ``ref_*`` functions restate the reference line by line (citations are
  ``file:line`` relative to the reference checkout).  They are PINNED: the
  golden vectors under ``tests/golden/`` were produced by importing the live
  reference (``tests/golden/make_golden.py``) and ``tests/test_oracle_golden.py``
  checks every ``ref_*`` function against them.  The reference itself ships no
  test, fixture or known-answer vector for this path, and its third-party
  arithmetic (OpenBLAS ``getrf/getri``/``potrf``/``gesv`` through Numba's
  ``np.linalg`` lowering, ``np.argsort``) is unpinned by the reference's own
  repo -- the golden vectors are outputs of the reference run in the build
  container (numba 0.65.0, numpy 2.3.5, scipy 1.18.1 OpenBLAS).
* ``chol_*`` functions are the better-conditioned Cholesky / ``W = L^-1``
  formulation the CUDA path uses.  They agree with ``ref_*`` to
  ``~eps * cond(K + 1e-6 I)`` (checked in the tests on well-conditioned inputs).
* ``exact_hvi_*`` (true 2-/3-objective hypervolume improvement) has NO reference
  implementation at all (the reference's "HVI" is sum-UCB): PARITY UNPINNED,
  the definition below is the specification.
"""

from __future__ import annotations

import os

import numpy as np

# Constants: reference bayesopt/config.py:54-66 (float64 build).
KERNEL_JITTER = 1e-6
CHOLESKY_JITTER = 1e-8
MIN_VARIANCE = 1e-10


# ----------------------------------------------------------------------------
# Restatement of the reference functions (explicit-inverse formulation)
# ----------------------------------------------------------------------------


_POOL = None
_TILE = 512


def _pool():
    """Host thread pool: NumPy ufuncs release the GIL, so column tiles run on all cores (the reference
    runs these loops under Numba ``prange``, numba_kernels.py:352, :432)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=os.cpu_count() or 1)
    return _POOL


def _sq_tile(a: np.ndarray, bt: np.ndarray, out: np.ndarray) -> None:
    """out[i, j] = sum_k (a[i,k] - bt[k,j])^2, accumulated k = 0..d-1 (direct difference form)."""
    tmp = np.empty_like(out)
    for k in range(a.shape[1]):
        np.subtract(a[:, k][:, None], bt[k][None, :], out=tmp)
        np.multiply(tmp, tmp, out=tmp)
        if k == 0:
            out[...] = tmp
        else:
            out += tmp


def _sq_dists(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Direct-difference squared distances, (len(a), len(b)).

    Reference: ``diff = x[i] - x[j]; sq = dot(diff, diff)``
    (numba_kernels.py:354-355 and :436-437).  Cache-blocked over column tiles on the host thread pool;
    the summation order is k = 0..d-1 for every pair.
    """
    a = np.ascontiguousarray(a, dtype=np.float64)
    bt = np.ascontiguousarray(np.asarray(b, dtype=np.float64).T)
    out = np.empty((a.shape[0], bt.shape[1]), dtype=np.float64)
    tiles = [(j, min(j + _TILE, bt.shape[1])) for j in range(0, bt.shape[1], _TILE)]
    if len(tiles) <= 1:
        for j0, j1 in tiles:
            _sq_tile(a, bt[:, j0:j1], out[:, j0:j1])
        return out
    list(_pool().map(lambda t: _sq_tile(a, bt[:, t[0]:t[1]], out[:, t[0]:t[1]]), tiles))
    return out


def ref_update_k(kernel_matrix, x_vector, last_eval, current_eval, prior_variance, length_scales):
    """RBF Gram matrix, in place.  Reference numba_kernels.py:329-367.

    Upper triangle of rows [last_eval, current_eval) then mirrored (:352-367).
    ``K[o,i,j] = var_o * exp(-0.5 * sq / ls_o**2)`` (:358-360).
    """
    n_obj = kernel_matrix.shape[0]
    rows = np.arange(last_eval, current_eval)
    if rows.size == 0:
        return
    sq = _sq_dists(x_vector[last_eval:current_eval], x_vector[last_eval:current_eval])
    iu = np.triu_indices(rows.size)
    for o in range(n_obj):
        block = prior_variance[o] * np.exp(-0.5 * sq / (length_scales[o] ** 2))
        upper = np.zeros_like(block)
        upper[iu] = block[iu]
        full = upper + np.triu(upper, 1).T
        kernel_matrix[o, last_eval:current_eval, last_eval:current_eval] = full


def ref_invert_k(current_eval, kernel_matrix):
    """``inv(K_o + 1e-6 I)`` per objective.  Reference numba_kernels.py:370-403."""
    n_obj = kernel_matrix.shape[0]
    out = np.zeros((n_obj, current_eval, current_eval), dtype=np.float64)
    for o in range(n_obj):
        k = np.array(kernel_matrix[o, :current_eval, :current_eval], dtype=np.float64)
        k[np.diag_indices(current_eval)] += KERNEL_JITTER  # :397-398
        out[o] = np.linalg.inv(k)  # :401
    return out


def ref_update_k_star(k_star, x_vector, input_space, last_eval, current_eval, prior_variance, length_scales):
    """Cross kernel, in place.  Reference numba_kernels.py:406-442 (one squared distance per pair, one exp
    per objective: :436-442)."""
    n_obj = k_star.shape[0]
    a = np.ascontiguousarray(x_vector[last_eval:current_eval], dtype=np.float64)
    bt = np.ascontiguousarray(np.asarray(input_space, dtype=np.float64).T)
    coef = [-0.5 / (length_scales[o] ** 2) for o in range(n_obj)]

    def work(t):
        j0, j1 = t
        sq = np.empty((a.shape[0], j1 - j0))
        _sq_tile(a, bt[:, j0:j1], sq)
        for o in range(n_obj):
            dst = k_star[o, last_eval:current_eval, j0:j1]
            np.multiply(sq, coef[o], out=dst)
            np.exp(dst, out=dst)
            dst *= prior_variance[o]

    tiles = [(j, min(j + _TILE, bt.shape[1])) for j in range(0, bt.shape[1], _TILE)]
    if len(tiles) <= 1:
        for t in tiles:
            work(t)
    else:
        list(_pool().map(work, tiles))


def ref_update_mean(mu_objectives, k_star, inverted_kernel_matrix, y_vector, prior_mean, current_eval):
    """Posterior mean, in place.  Reference numba_kernels.py:450-488."""
    n = current_eval
    for o in range(mu_objectives.shape[0]):
        kinv = np.ascontiguousarray(inverted_kernel_matrix[o, :n, :n])
        delta = np.ascontiguousarray(y_vector[:n, o] - prior_mean[o])  # :477-479
        partial = kinv @ delta  # :483
        # :486-488 (the reference first makes a contiguous copy of k_star.T; the product is the same gemv)
        mu_objectives[o, :] = prior_mean[o] + k_star[o, :n, :].T @ partial


def ref_update_variance(variance_objectives, k_star, inverted_kernel_matrix, prior_variance, current_eval):
    """Posterior variance with the absolute 1e-10 clamp.  Reference numba_kernels.py:491-535."""
    n = current_eval
    for o in range(variance_objectives.shape[0]):
        kinv = np.ascontiguousarray(inverted_kernel_matrix[o, :n, :n])
        ks = np.ascontiguousarray(k_star[o, :n, :])
        inter = kinv @ ks  # :521
        quad = np.einsum("ij,ij->j", ks, inter)  # :525-529 (column dot products)
        variance_objectives[o, :] = np.maximum(prior_variance[o] - quad, MIN_VARIANCE)  # :532-535


def ref_standardize_objectives(std_mu, std_var, mu, var, prior_mean, prior_variance):
    """Reference numba_kernels.py:538-570."""
    for o in range(mu.shape[0]):
        std_mu[o] = (mu[o] - prior_mean[o]) / np.sqrt(prior_variance[o])  # :563-565
        std_var[o] = var[o] / prior_variance[o]  # :568-570


def ref_upper_confidence_bound(mu, variance, beta):
    """Reference acquisition.py:33-52."""
    return mu + beta * np.sqrt(np.abs(variance))


def ref_update_ucb(ucb, mu_objectives, variance_objectives, betas):
    """Reference acquisition.py:55-81."""
    for o in range(mu_objectives.shape[0]):
        ucb[o] = ref_upper_confidence_bound(mu_objectives[o], variance_objectives[o], betas[o])


def ref_update_hypervolume_improvement(acquisition_values, ucb):
    """Sum of UCB over objectives, sequential from 0.0.  Reference acquisition.py:89-108.

    ``np.sum(ucb[:, i])`` over a length-m strided column adds left to right
    starting from 0.0 (SURVEY: (1e16, 1, -1e16) -> 0.0).
    """
    acc = np.zeros(ucb.shape[1], dtype=np.float64)
    for o in range(ucb.shape[0]):
        acc = acc + ucb[o]
    acquisition_values[:] = acc


def ranked_indices(acquisition_values: np.ndarray) -> np.ndarray:
    """Descending order with the tie rule of the CUDA path: value desc, index asc.

    The reference uses ``np.argsort(acq)[::-1]`` (acquisition.py:134) whose tie
    order is unspecified; on tie-free inputs both give the same permutation.
    NaNs sort last here (np.argsort puts them last ascending, i.e. FIRST after
    the reversal in the reference -- inputs with NaN scores are outside the
    parity contract).
    """
    a = np.asarray(acquisition_values, dtype=np.float64)
    key = np.where(np.isnan(a), -np.inf, a)
    return np.lexsort((np.arange(a.size), -key))


def ref_select_next_batch(input_space, acquisition_values, evaluated_points, batch_size=3):
    """Reference acquisition.py:116-144 with the deterministic tie rule above.

    Walk candidates best->worst, skip rows exactly equal to an evaluated row
    (:139), stop at ``batch_size``.  Returns ``np.array(batch)`` exactly like
    the reference (dtype of ``input_space``; possibly fewer rows).
    Also returns the selected indices as a second value.
    """
    order = ranked_indices(acquisition_values)
    batch, picked = [], []
    ev = np.asarray(evaluated_points)
    for idx in order:
        cand = input_space[idx]
        if ev.shape[0] == 0 or not np.any(np.all(cand == ev, axis=1)):
            batch.append(cand)
            picked.append(int(idx))
            if len(batch) == batch_size:
                break
    return np.array(batch), np.array(picked, dtype=np.int64)


def ref_is_pareto_efficient_loop(y_vector: np.ndarray) -> np.ndarray:
    """Literal restatement of the reference's skip/break loop.  Reference pareto.py:12-45."""
    yn = -np.asarray(y_vector)
    n = yn.shape[0]
    eff = np.ones(n, dtype=bool)
    for i in range(n):
        if not eff[i]:
            continue
        for j in range(i + 1, n):
            if np.all(yn[j] <= yn[i]) and np.any(yn[j] < yn[i]):
                eff[i] = False
                break
            if np.all(yn[i] <= yn[j]) and np.any(yn[i] < yn[j]):
                eff[j] = False
    return eff


def pareto_mask_definition(y_vector: np.ndarray, block: int = 1024) -> np.ndarray:
    """Order-free definition the loop above is equivalent to (maximisation).

    i is dropped iff some j has ``all(y_j >= y_i) and any(y_j > y_i)``;
    duplicates both stay, rows containing NaN never dominate nor are dominated.
    Blocked O(n^2) so that n ~ 1e5 stays tractable on the CPU.
    """
    y = np.asarray(y_vector, dtype=np.float64)
    n = y.shape[0]
    eff = np.ones(n, dtype=bool)
    for i0 in range(0, n, block):
        yi = y[i0 : i0 + block]  # (bi, m)
        dominated = np.zeros(yi.shape[0], dtype=bool)
        for j0 in range(0, n, block):
            yj = y[j0 : j0 + block]  # (bj, m)
            ge = np.all(yj[None, :, :] >= yi[:, None, :], axis=2)
            gt = np.any(yj[None, :, :] > yi[:, None, :], axis=2)
            dominated |= np.any(ge & gt, axis=1)
        eff[i0 : i0 + block] = ~dominated
    return eff


def ref_compute_mll(x_vector, y_vector, kernel_matrix, prior_mean, prior_variance, length_scales, current_eval,
                    jitter=CHOLESKY_JITTER):
    """Log marginal likelihood summed over objectives.  Reference numba_kernels.py:152-235.

    Side effect kept: ``kernel_matrix`` is overwritten by ``update_k`` (:178).
    """
    ref_update_k(kernel_matrix, x_vector, 0, current_eval, prior_variance, length_scales)
    n = current_eval
    total = []
    for o in range(y_vector.shape[1]):
        k = np.ascontiguousarray(kernel_matrix[o, :n, :n] / prior_variance[o])  # :195-198
        yc = np.ascontiguousarray(y_vector[:n, o] - prior_mean[o])  # :201-204
        s = np.std(yc)  # population std, :206
        if s > 0.0:
            yc = yc / s
        l = np.linalg.cholesky(k + jitter * np.eye(n))  # :211-214
        inter = np.linalg.solve(l, yc)  # :216
        alpha = np.linalg.solve(l.T, inter)  # :219
        fit = -0.5 * np.dot(yc, alpha)  # :222
        logdet = 2.0 * np.sum(np.log(np.diag(l)))  # :225
        total.append(fit - 0.5 * logdet - 0.5 * n * np.log(2.0 * np.pi))  # :226-232
    return float(np.sum(np.array(total)))  # :235


def ref_hot_path(x_vector, y_vector, input_space, prior_mean, prior_variance, length_scales, betas,
                 current_eval, batch_size=3):
    """The per-iteration sequence of the reference loop, steps b..h.

    Reference bayesian_optimization.py:129-207.  Materialises ``k_star`` exactly
    like the reference does, so use it on candidate chunks only.
    Returns a dict with every array the reference's ``state`` exposes.
    """
    n = current_eval
    m = y_vector.shape[1]
    n_cand = input_space.shape[0]
    kmat = np.zeros((m, n, n))
    ref_update_k(kmat, x_vector, 0, n, prior_variance, length_scales)
    kinv = ref_invert_k(n, kmat)
    k_star = np.zeros((m, n, n_cand))
    ref_update_k_star(k_star, x_vector, input_space, 0, n, prior_variance, length_scales)
    mu = np.zeros((m, n_cand))
    var = np.zeros((m, n_cand))
    ref_update_mean(mu, k_star, kinv, y_vector, prior_mean, n)
    ref_update_variance(var, k_star, kinv, prior_variance, n)
    smu, svar, ucb = np.zeros_like(mu), np.zeros_like(var), np.zeros_like(mu)
    ref_standardize_objectives(smu, svar, mu, var, prior_mean, prior_variance)
    ref_update_ucb(ucb, smu, svar, betas)
    acq = np.zeros(n_cand)
    ref_update_hypervolume_improvement(acq, ucb)
    x_next, idx = ref_select_next_batch(input_space, acq, x_vector[:n], batch_size)
    return dict(kernel=kmat, kinv=kinv, mu=mu, var=var, std_mu=smu, std_var=svar, ucb=ucb, acq=acq,
                x_next=x_next, idx=idx)


# ----------------------------------------------------------------------------
# Cholesky / W = L^-1 formulation (what the CUDA path computes)
# ----------------------------------------------------------------------------


def chol_fit(x_vector, y_vector, prior_mean, prior_variance, length_scales, current_eval, jitter=KERNEL_JITTER):
    """Factor ``K_o + jitter I = L L^T``; return ``L``, ``W = L^-1`` and ``alpha = K^-1 (y - mu0)``."""
    n = current_eval
    m = y_vector.shape[1]
    kmat = np.zeros((m, n, n))
    ref_update_k(kmat, x_vector, 0, n, prior_variance, length_scales)
    L = np.zeros_like(kmat)
    W = np.zeros_like(kmat)
    alpha = np.zeros((m, n))
    eye = np.eye(n)
    for o in range(m):
        L[o] = np.linalg.cholesky(kmat[o] + jitter * eye)
        W[o] = np.linalg.solve(L[o], eye)  # lower triangular inverse
        W[o] = np.tril(W[o])
        alpha[o] = W[o].T @ (W[o] @ (y_vector[:n, o] - prior_mean[o]))
    return dict(kernel=kmat, L=L, W=W, alpha=alpha)


def chol_predict(fit, x_vector, input_space, prior_mean, prior_variance, length_scales, current_eval):
    """``mu = mu0 + k*^T alpha``; ``var = max(var0 - ||W k*||^2, 1e-10)``."""
    n = current_eval
    m = fit["W"].shape[0]
    sq = _sq_dists(x_vector[:n], input_space)
    mu = np.zeros((m, input_space.shape[0]))
    var = np.zeros_like(mu)
    for o in range(m):
        ks = prior_variance[o] * np.exp(-0.5 * sq / (length_scales[o] ** 2))
        mu[o] = prior_mean[o] + ks.T @ fit["alpha"][o]
        v = fit["W"][o] @ ks
        var[o] = np.maximum(prior_variance[o] - np.einsum("ij,ij->j", v, v), MIN_VARIANCE)
    return mu, var


def chol_hot_path(x_vector, y_vector, input_space, prior_mean, prior_variance, length_scales, betas,
                  current_eval, batch_size=3):
    """Cholesky-form counterpart of :func:`ref_hot_path` (same outputs)."""
    fit = chol_fit(x_vector, y_vector, prior_mean, prior_variance, length_scales, current_eval)
    mu, var = chol_predict(fit, x_vector, input_space, prior_mean, prior_variance, length_scales, current_eval)
    smu, svar, ucb = np.zeros_like(mu), np.zeros_like(var), np.zeros_like(mu)
    ref_standardize_objectives(smu, svar, mu, var, prior_mean, prior_variance)
    ref_update_ucb(ucb, smu, svar, betas)
    acq = np.zeros(mu.shape[1])
    ref_update_hypervolume_improvement(acq, ucb)
    x_next, idx = ref_select_next_batch(input_space, acq, x_vector[:current_eval], batch_size)
    return dict(mu=mu, var=var, std_mu=smu, std_var=svar, ucb=ucb, acq=acq, x_next=x_next, idx=idx, **fit)


def mll_grid(x_vector, y_vector, prior_mean, length_scale_grid, jitter_grid, current_eval):
    """cfg5: LML for every (length-scale, jitter) setting; all objectives share the setting.

    Equals :func:`ref_compute_mll` at ``jitter == CHOLESKY_JITTER`` (the MLL does
    not depend on prior_variance because the Gram matrix is normalised, :195-197).
    Returns (S,) float64 with S = len(length_scale_grid) = len(jitter_grid).
    """
    n = current_eval
    m = y_vector.shape[1]
    out = np.zeros(len(length_scale_grid))
    ones = np.ones(m)
    for s, (ls, jit) in enumerate(zip(length_scale_grid, jitter_grid)):
        kmat = np.zeros((m, n, n))
        out[s] = ref_compute_mll(x_vector, y_vector, kmat, prior_mean, ones, np.full(m, ls), n, jitter=jit)
    return out


# ----------------------------------------------------------------------------
# Exact hypervolume improvement (opt-in mode; no reference implementation)
# ----------------------------------------------------------------------------


def hypervolume_2d(front: np.ndarray, ref: np.ndarray) -> float:
    """Area dominated by ``front`` (maximisation) above ``ref``.  Points not above ref are clipped."""
    f = np.maximum(np.asarray(front, dtype=np.float64).reshape(-1, 2), ref)
    if f.shape[0] == 0:
        return 0.0
    order = np.lexsort((-f[:, 1], -f[:, 0]))  # objective 0 descending
    hv, best1 = 0.0, ref[1]
    for a, b in f[order]:
        if b > best1:
            hv += (a - ref[0]) * (b - best1)
            best1 = b
    return float(hv)


def hypervolume_3d(front: np.ndarray, ref: np.ndarray) -> float:
    """Volume dominated by ``front`` above ``ref``: slice on objective 2, 2-D area per slab."""
    f = np.maximum(np.asarray(front, dtype=np.float64).reshape(-1, 3), ref)
    if f.shape[0] == 0:
        return 0.0
    order = np.argsort(-f[:, 2], kind="stable")
    f = f[order]
    hv = 0.0
    for i in range(f.shape[0]):
        z_hi = f[i, 2]
        z_lo = f[i + 1, 2] if i + 1 < f.shape[0] else ref[2]
        if z_hi > z_lo:
            hv += hypervolume_2d(f[: i + 1, :2], ref[:2]) * (z_hi - z_lo)
    return float(hv)


def exact_hvi(points: np.ndarray, front: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """``HVI(u) = HV(front U {u}) - HV(front)`` for each row u of ``points`` (m = 2 or 3).

    Specification of the opt-in exact mode (discrepancy 1).  PARITY
    UNPINNED: the reference has nothing to compare with.
    """
    points = np.asarray(points, dtype=np.float64)
    front = np.asarray(front, dtype=np.float64).reshape(-1, points.shape[1])
    ref = np.asarray(ref, dtype=np.float64)
    hv = hypervolume_2d if points.shape[1] == 2 else hypervolume_3d
    base = hv(front, ref)
    out = np.zeros(points.shape[0])
    for i, u in enumerate(points):
        out[i] = hv(np.vstack([front, u[None, :]]), ref) - base
    return out


# ----------------------------------------------------------------------------
# Synthetic workloads named in BASELINE.json; deterministic.
# ----------------------------------------------------------------------------


def zdt1(x: np.ndarray) -> np.ndarray:
    g = 1.0 + 9.0 * np.mean(x[:, 1:], axis=1)
    f1 = x[:, 0]
    return np.stack([f1, g * (1.0 - np.sqrt(f1 / g))], axis=1)


def zdt2(x: np.ndarray) -> np.ndarray:
    g = 1.0 + 9.0 * np.mean(x[:, 1:], axis=1)
    f1 = x[:, 0]
    return np.stack([f1, g * (1.0 - (f1 / g) ** 2)], axis=1)


def dtlz2(x: np.ndarray, n_obj: int = 3) -> np.ndarray:
    g = np.sum((x[:, n_obj - 1 :] - 0.5) ** 2, axis=1)
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
    """X ~ U[0,1]^{n x d}; y = -objective (reference maximises); mu0 = mean, var0 = var."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    fn = {"zdt1": zdt1, "zdt2": zdt2, "dtlz2": dtlz2}[name]
    y = -fn(x)
    return x, y, y.mean(axis=0), y.var(axis=0)
