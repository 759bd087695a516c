"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (the reference checkout does not travel to the
GPU box):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden.py

Every array written here is an output of the unmodified reference functions
(alebal123bal/BayesOpt_smart, Numba mode: numba 0.65.0, numpy 2.3.5, scipy
1.18.1 / OpenBLAS).  The committed ``*.npz`` files are what
``tests/test_oracle_golden.py`` pins ``oracle/gp_oracle.py`` against and what the
``-m gpu`` parity tests compare the CUDA path with.
"""

import os
import sys

import numpy as np

REF = os.environ.get("BAYESOPT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from bayesopt import numba_kernels as nk  # noqa: E402
from bayesopt import acquisition as acq_mod  # noqa: E402
from bayesopt import pareto as pareto_mod  # noqa: E402
from bayesopt.bayesian_optimization import BayesianOptimization  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference_path(x_vector, y_vector, input_space, prior_mean, prior_variance, length_scales, betas, n,
                       batch_size):
    """Steps b..h of the reference loop (bayesian_optimization.py:129-207), unmodified functions."""
    total, m = y_vector.shape
    n_cand = input_space.shape[0]
    kmat = np.zeros((m, total, total))
    nk.update_k(kernel_matrix=kmat, x_vector=x_vector, last_eval=0, current_eval=n,
                prior_variance=prior_variance, length_scales=length_scales)
    kinv = nk.invert_k(current_eval=n, kernel_matrix=kmat)
    k_star = np.zeros((m, total, n_cand))
    nk.update_k_star(k_star=k_star, x_vector=x_vector, input_space=input_space, last_eval=0, current_eval=n,
                     prior_variance=prior_variance, length_scales=length_scales)
    mu = np.zeros((m, n_cand))
    var = np.zeros((m, n_cand))
    nk.update_mean(mu_objectives=mu, k_star=k_star, inverted_kernel_matrix=kinv, y_vector=y_vector,
                   prior_mean=prior_mean, current_eval=n)
    nk.update_variance(variance_objectives=var, k_star=k_star, inverted_kernel_matrix=kinv,
                       prior_variance=prior_variance, current_eval=n)
    smu, svar, ucb = np.zeros_like(mu), np.zeros_like(mu), np.zeros_like(mu)
    nk.standardize_objectives(std_mu_objectives=smu, std_variance_objectives=svar, mu_objectives=mu,
                              variance_objectives=var, prior_mean=prior_mean, prior_variance=prior_variance)
    acq_mod.update_ucb(ucb=ucb, mu_objectives=smu, variance_objectives=svar, betas=betas)
    acq = np.zeros(n_cand)
    acq_mod.update_hypervolume_improvement(acquisition_values=acq, ucb=ucb)
    x_next = acq_mod.select_next_batch(input_space=input_space, acquisition_values=acq,
                                       evaluated_points=x_vector[:n], batch_size=batch_size)
    return dict(kernel=kmat[:, :n, :n].copy(), kinv=kinv, k_star=k_star[:, :n, :].copy(), mu=mu, var=var,
                std_mu=smu, std_var=svar, ucb=ucb, acq=acq, x_next=x_next,
                cond=np.array([np.linalg.cond(kmat[o, :n, :n] + 1e-6 * np.eye(n)) for o in range(m)]))


def case_float(seed, n, total, d, m, n_cand, ls, betas, batch_size):
    rng = np.random.default_rng(seed)
    x = np.zeros((total, d))
    x[:n] = rng.random((n, d))
    w = rng.normal(size=(d, m))
    y = np.zeros((total, m))
    y[:n] = np.sin(3.0 * x[:n] @ w) + 0.1 * (x[:n] ** 2).sum(axis=1, keepdims=True)
    cand = rng.random((n_cand, d))
    cand[5] = x[2]  # one candidate coincides with a training point (exclusion + clamp region)
    mu0 = y[:n].mean(axis=0)
    var0 = y[:n].var(axis=0)
    ls = np.array(ls, dtype=np.float64)
    betas = np.array(betas, dtype=np.float64)
    out = run_reference_path(x, y, cand, mu0, var0, ls, betas, n, batch_size)
    out.update(x_vector=x, y_vector=y, input_space=cand, prior_mean=mu0, prior_variance=var0, length_scales=ls,
               betas=betas, n=np.int64(n), batch_size=np.int64(batch_size))
    return out


def case_int_grid(n, m, batch_size):
    """int64 Cartesian grid like bayesian_optimization.py:338-340; evaluated points are grid members."""
    rng = np.random.default_rng(7)
    ranges = [np.arange(0, 12), np.arange(0, 12)]
    mesh = np.meshgrid(*ranges, indexing="ij")
    cand = np.stack([g.ravel() for g in mesh], axis=-1)
    assert cand.dtype == np.int64
    total = n + 6
    pick = rng.choice(cand.shape[0], size=n, replace=False)
    x = np.zeros((total, 2))
    x[:n] = cand[pick]
    centres = np.array([[3.0, 8.0], [9.0, 2.0], [6.0, 6.0]])[:m]
    y = np.zeros((total, m))
    for o in range(m):
        y[:n, o] = -((x[:n] - centres[o]) ** 2).sum(axis=1) + 10.0 * (o + 1)
    mu0 = y[:n].mean(axis=0)
    var0 = y[:n].var(axis=0)
    ls = np.array([2.0, 2.5, 3.0])[:m]
    betas = np.array([1.0, 2.0, 0.5])[:m]
    out = run_reference_path(x, y, cand, mu0, var0, ls, betas, n, batch_size)
    out.update(x_vector=x, y_vector=y, input_space=cand, prior_mean=mu0, prior_variance=var0, length_scales=ls,
               betas=betas, n=np.int64(n), batch_size=np.int64(batch_size))
    return out


def case_mll():
    rng = np.random.default_rng(11)
    n, total, d, m = 40, 48, 4, 2
    x = np.zeros((total, d))
    x[:n] = rng.random((n, d))
    y = np.zeros((total, m))
    y[:n, 0] = np.sin(4 * x[:n, 0]) + x[:n, 1]
    y[:n, 1] = np.cos(3 * x[:n, 2]) * x[:n, 3]
    mu0 = y[:n].mean(axis=0)
    settings = np.array([[0.2, 0.2, 1.0, 1.0], [0.5, 0.7, 1.0, 3.0], [1.0, 0.3, 2.5e3, 0.1], [2.0, 2.0, 1.0, 1.0]])
    vals = []
    for s in settings:
        kmat = np.zeros((m, total, total))
        vals.append(nk.compute_mll(x_vector=x, y_vector=y, kernel_matrix=kmat, prior_mean=mu0,
                                   prior_variance=s[2:].copy(), length_scales=s[:2].copy(), current_eval=n))
    return dict(x_vector=x, y_vector=y, prior_mean=mu0, settings=settings, mll=np.array(vals), n=np.int64(n))


def case_pareto():
    rng = np.random.default_rng(3)
    sets = {}
    sets["kat"] = np.array([[1, 1], [1, 1], [0, 2], [0, 0], [np.nan, 5]], dtype=np.float64)
    sets["rand2"] = rng.normal(size=(96, 2))
    y3 = rng.normal(size=(120, 3))
    y3[10] = y3[3]  # duplicate rows both kept
    y3[20, 1] = np.nan
    sets["rand3_dup_nan"] = y3
    sets["ints2"] = rng.integers(0, 6, size=(150, 2)).astype(np.float64)  # many ties
    t = np.linspace(0, 1, 64)
    sets["all_efficient"] = np.stack([t, 1 - t], axis=1)
    sets["single"] = np.array([[2.0, 3.0, 4.0]])
    out = {}
    for k, v in sets.items():
        out[k + "_y"] = v
        out[k + "_mask"] = pareto_mod.is_pareto_efficient(v)
    return out


def case_cfg1_trace():
    """BASELINE config 1: headless demo, toy_function, initial_samples=10, n_iterations=20 (batch 3)."""
    sys.path.insert(0, REF)
    from examples.benchmark_functions import toy_function

    trace = []

    def grab(state):
        it = len(trace)
        rec = dict(iteration=state["iteration"], hyperparams=np.array(state["hyperparams"], dtype=np.float64),
                   x_next=np.array(state["x_next"]), n_evaluations=state["n_evaluations"])
        if it < 2:
            rec["mu"] = state["mu_objectives"].copy()
            rec["var"] = state["variance_objectives"].copy()
            rec["acq"] = state["acquisition_values"].copy()
        trace.append(rec)

    opt = BayesianOptimization(function=toy_function, bounds=[(0, 300), (0, 300)], n_objectives=2, n_iterations=20,
                               initial_samples=10, callbacks=[grab])
    x0 = opt.x_vector[:10].copy()
    y0 = opt.y_vector[:10].copy()
    mu0 = opt.prior_mean.copy()
    var0 = opt.prior_variance.copy()
    opt.optimize()
    front = opt.pareto_analysis()
    sub = np.arange(0, 90000, 97)
    out = dict(x_init=x0, y_init=y0, prior_mean=mu0, prior_variance_init=var0, betas=opt.betas.copy(),
               x_vector=opt.x_vector.copy(), y_vector=opt.y_vector.copy(), n_evaluations=np.int64(opt.n_evaluations),
               pareto_front=front, sub_index=sub,
               hyperparams=np.stack([t["hyperparams"] for t in trace]),
               x_next=np.stack([t["x_next"] for t in trace]),
               iteration=np.array([t["iteration"] for t in trace]))
    for it in range(2):
        out[f"mu_sub_{it}"] = trace[it]["mu"][:, sub]
        out[f"var_sub_{it}"] = trace[it]["var"][:, sub]
        out[f"acq_sub_{it}"] = trace[it]["acq"][sub]
        order = np.argsort(-trace[it]["acq"], kind="stable")[:16]
        out[f"top16_idx_{it}"] = order
        out[f"top16_val_{it}"] = trace[it]["acq"][order]
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "gp_float_m2.npz"),
                        **case_float(seed=0, n=24, total=30, d=3, m=2, n_cand=257, ls=[0.6, 0.8], betas=[2.0, 1.5],
                                     batch_size=3))
    np.savez_compressed(os.path.join(HERE, "gp_float_m3.npz"),
                        **case_float(seed=1, n=96, total=100, d=6, m=3, n_cand=300, ls=[0.5, 0.6, 0.7],
                                     betas=[2.0, 2.0, 1.0], batch_size=4))
    np.savez_compressed(os.path.join(HERE, "gp_intgrid_m3.npz"), **case_int_grid(n=9, m=3, batch_size=5))
    np.savez_compressed(os.path.join(HERE, "mll.npz"), **case_mll())
    np.savez_compressed(os.path.join(HERE, "pareto.npz"), **case_pareto())
    if os.environ.get("GOLDEN_SKIP_CFG1", "0") != "1":
        np.savez_compressed(os.path.join(HERE, "cfg1_trace.npz"), **case_cfg1_trace())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
