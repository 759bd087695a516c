"""Drop-in for the reference's ``bayesopt/bayesian_optimization.py``: same class, constructor
kwargs, public attributes, ``optimize()`` / ``pareto_analysis()`` and callback contract; the
per-iteration hot path runs device-resident through :class:`engine.DeviceGP`.
"""
from __future__ import annotations

import time
from typing import Any, Callable, List, Optional, Tuple

import numpy as np
import torch

from .acquisition import exact_hvi_device
from .config import (
    DEFAULT_BATCH_SIZE,
    DEFAULT_BETA,
    DEFAULT_INITIAL_SAMPLES,
    DEFAULT_LENGTH_SCALE,
    DEFAULT_PRIOR_MEAN,
    DEFAULT_PRIOR_VARIANCE,
    NUMBA_FLOAT_TYPE,
)
from .engine import DeviceGP, PinnedMirror, require_cuda, to_device
from .numba_kernels import (
    compute_prior_mean,
    compute_prior_variance,
    initialize_lhs_integer,
    optimize_hyperparams_mll,
)
from .pareto import compute_pareto_front, is_pareto_efficient, print_pareto_analysis


def optimize(
    x_vector: np.ndarray,
    y_vector: np.ndarray,
    kernel_matrices: np.ndarray,
    k_star: np.ndarray,
    mu_objectives: np.ndarray,
    variance_objectives: np.ndarray,
    std_mu_objectives: np.ndarray,
    std_variance_objectives: np.ndarray,
    ucb: np.ndarray,
    acquisition_values: np.ndarray,
    input_space: np.ndarray,
    prior_mean: np.ndarray,
    prior_variance: np.ndarray,
    reference_point: np.ndarray,
    n_evaluations: int,
    total_samples: int,
    n_objectives: int,  # pylint: disable=unused-argument
    function: Callable[[np.ndarray], np.ndarray],
    betas: np.ndarray,
    length_scales: np.ndarray,
    batch_size: int,
    bounds: List[Tuple[int, int]],  # pylint: disable=unused-argument
    callbacks: Optional[List[Callable]] = None,
    acquisition: str = "sum_ucb",
) -> Tuple[np.ndarray, np.ndarray, int]:
    """The BO loop.  Reference bayesian_optimization.py:51-247 (same parameters, same return).

    Per iteration: Powell fit of the hyper-parameters on the GPU MLL, ``DeviceGP.fit`` (Gram +
    Cholesky + W = L^-1 + alpha), ``DeviceGP.score`` over the resident candidate set,
    ``DeviceGP.select``.  The host arrays ``mu_objectives`` .. ``acquisition_values`` are refreshed
    every iteration when callbacks are installed (they receive NumPy arrays) and once at the end
    otherwise.  ``kernel_matrices`` and ``k_star`` are accepted for signature compatibility and left
    untouched: K* is never materialised.  ``acquisition="exact_hvi"`` (opt-in, 2 or 3 objectives)
    replaces the reference's sum-UCB score by the exact hypervolume improvement of the UCB vector
    against the current standardised front, with ``reference_point`` as lower corner.
    Returns ``(x_vector, y_vector, last_eval + 1)`` like the reference (:247).
    """
    dev = require_cuda()
    gp = DeviceGP(dev)
    cand_dev = to_device(input_space, None, dev)  # uploaded once, resident for the whole run
    n_cand = cand_dev.shape[0]
    m = y_vector.shape[1]
    keys = ("mu", "var", "std_mu", "std_var", "ucb", "acq")
    out = {k: torch.empty((n_cand,) if k == "acq" else (m, n_cand), dtype=torch.float64, device=dev) for k in keys}
    host = dict(mu=mu_objectives, var=variance_objectives, std_mu=std_mu_objectives, std_var=std_variance_objectives,
                ucb=ucb, acq=acquisition_values)
    mirror = PinnedMirror()
    last_eval = 0
    iterations = list(range(n_evaluations, total_samples, batch_size))
    for current_eval in iterations:
        iter_start = time.perf_counter()
        t0 = time.perf_counter()
        optimized_hyperparams = optimize_hyperparams_mll(
            x_vector=x_vector, y_vector=y_vector, kernel_matrix=kernel_matrices, prior_mean=prior_mean,
            prior_variance=prior_variance, length_scales=length_scales, current_eval=current_eval)
        t1 = time.perf_counter()

        gp.fit(x_vector[:current_eval], y_vector[:current_eval], prior_mean, prior_variance, length_scales,
               current_eval)
        torch.cuda.synchronize()
        t2 = time.perf_counter()

        gp.score(cand_dev, betas, out=out)
        score = out["acq"]
        if acquisition == "exact_hvi":
            y_std = (y_vector[:current_eval] - prior_mean) / np.sqrt(prior_variance)
            front = y_std[is_pareto_efficient(y_std)]
            ref_std = (np.asarray(reference_point, dtype=np.float64) - prior_mean) / np.sqrt(prior_variance)
            ref_std = np.minimum(ref_std, y_std.min(axis=0))
            score = exact_hvi_device(out["ucb"], front, ref_std)
            out["acq"].copy_(score)
        ev_dev = to_device(x_vector[:current_eval], torch.float64, dev)
        _, idx = gp.select(cand_dev, score, ev_dev, batch_size)
        x_next = np.array([input_space[i] for i in idx])
        is_last = current_eval == iterations[-1]
        if callbacks or is_last:
            staged = {k: mirror.get(k, tuple(out[k].shape)) for k in keys}
            for k in keys:
                staged[k].copy_(out[k], non_blocking=True)
            torch.cuda.synchronize()
            for k in keys:
                host[k][...] = staged[k].numpy()
        t3 = time.perf_counter()

        for b_idx, point in enumerate(x_next):
            x_vector[current_eval + b_idx] = point
            y_vector[current_eval + b_idx] = function(point)
        last_eval = current_eval
        t4 = time.perf_counter()

        if callbacks:
            state = {
                "iteration": current_eval,
                "n_evaluations": current_eval + batch_size,
                "x_vector": x_vector[: current_eval + batch_size],
                "y_vector": y_vector[: current_eval + batch_size],
                "mu_objectives": mu_objectives,
                "variance_objectives": variance_objectives,
                "acquisition_values": acquisition_values,
                "x_next": x_next,
                "hyperparams": optimized_hyperparams.x,
                "timings": {
                    "hyperparams": t1 - t0,
                    "kernels": t2 - t1,
                    "acquisition": t3 - t2,
                    "eval": t4 - t3,
                    "total": t4 - iter_start,
                },
            }
            for callback in callbacks:
                callback(state)

    return x_vector, y_vector, last_eval + 1


class BayesianOptimization:
    """Multi-objective Bayesian optimisation; reference bayesian_optimization.py:250-488."""

    def __init__(self, function: Callable[[np.ndarray], np.ndarray], bounds: List[Tuple[int, int]],
                 n_objectives: int = 3, n_iterations: int = 10, **kwargs: Any):
        """Same arguments and kwargs as the reference (:259-332): callbacks, prior_mean, prior_variance,
        length_scales, betas, batch_size, initial_samples.  Extra opt-in kwarg: ``acquisition``
        ("sum_ucb" default = reference behaviour, or "exact_hvi")."""
        self.function = function
        self.bounds = bounds
        self.n_objectives = n_objectives
        self.n_iterations = n_iterations

        callbacks_param = kwargs.get("callbacks", None)
        if callbacks_param is not None:
            self.callbacks = callbacks_param if isinstance(callbacks_param, list) else [callbacks_param]
        else:
            self.callbacks = []

        self.prior_mean = np.array(kwargs.get("prior_mean", [DEFAULT_PRIOR_MEAN] * n_objectives),
                                   dtype=NUMBA_FLOAT_TYPE)
        self.prior_variance = np.array(kwargs.get("prior_variance", [DEFAULT_PRIOR_VARIANCE] * n_objectives),
                                       dtype=NUMBA_FLOAT_TYPE)
        self.length_scales = np.array(kwargs.get("length_scales", [DEFAULT_LENGTH_SCALE] * n_objectives),
                                      dtype=NUMBA_FLOAT_TYPE)
        self.betas = np.array(kwargs.get("betas", [DEFAULT_BETA] * n_objectives), dtype=NUMBA_FLOAT_TYPE)
        self.batch_size = kwargs.get("batch_size", DEFAULT_BATCH_SIZE)
        self.initial_samples = kwargs.get("initial_samples", DEFAULT_INITIAL_SAMPLES)
        self.acquisition = kwargs.get("acquisition", "sum_ucb")
        self.dim = len(bounds)

        # integer Cartesian grid, upper bound exclusive (:338-340)
        ranges = [np.arange(b[0], b[1]) for b in bounds]
        mesh = np.meshgrid(*ranges, indexing="ij")
        self.input_space = np.stack([g.ravel() for g in mesh], axis=-1)
        n_cand = len(self.input_space)

        self.total_samples = self.initial_samples + self.n_iterations * self.batch_size
        self.x_vector = np.zeros((self.total_samples, self.dim), dtype=NUMBA_FLOAT_TYPE)
        self.y_vector = np.zeros((self.total_samples, n_objectives), dtype=NUMBA_FLOAT_TYPE)
        self.kernel_matrices = np.zeros((n_objectives, self.total_samples, self.total_samples),
                                        dtype=NUMBA_FLOAT_TYPE)
        # the reference preallocates k_star (m, T, M) here (:362-365); the fused path never materialises it
        self.k_star = None
        self.mu_objectives = np.zeros((n_objectives, n_cand), dtype=NUMBA_FLOAT_TYPE)
        self.variance_objectives = np.zeros((n_objectives, n_cand), dtype=NUMBA_FLOAT_TYPE)
        self.std_mu_objectives = np.zeros((n_objectives, n_cand), dtype=NUMBA_FLOAT_TYPE)
        self.std_variance_objectives = np.zeros((n_objectives, n_cand), dtype=NUMBA_FLOAT_TYPE)
        self.ucb = np.zeros((n_objectives, n_cand), dtype=NUMBA_FLOAT_TYPE)
        self.acquisition_values = np.zeros(n_cand, dtype=NUMBA_FLOAT_TYPE)

        self.n_evaluations = initialize_lhs_integer(
            x_vector=self.x_vector, y_vector=self.y_vector, bounds=np.array(self.bounds, dtype=np.int64),
            function=self.function, n_samples=self.initial_samples)

        if np.all(self.prior_mean == DEFAULT_PRIOR_MEAN):  # exact-equality auto-detection (:413)
            self.prior_mean = compute_prior_mean(self.y_vector, self.n_evaluations, n_objectives)
        if np.all(self.prior_variance == DEFAULT_PRIOR_VARIANCE):  # (:419)
            self.prior_variance = compute_prior_variance(self.y_vector, self.n_evaluations, n_objectives)

        self.reference_point = np.array([0.0] * n_objectives)

    def optimize(self) -> None:
        """Run the optimisation loop (:427-463)."""
        self.x_vector, self.y_vector, self.n_evaluations = optimize(
            x_vector=self.x_vector, y_vector=self.y_vector, kernel_matrices=self.kernel_matrices, k_star=self.k_star,
            mu_objectives=self.mu_objectives, variance_objectives=self.variance_objectives,
            std_mu_objectives=self.std_mu_objectives, std_variance_objectives=self.std_variance_objectives,
            ucb=self.ucb, acquisition_values=self.acquisition_values, input_space=self.input_space,
            prior_mean=self.prior_mean, prior_variance=self.prior_variance, reference_point=self.reference_point,
            n_evaluations=self.n_evaluations, total_samples=self.total_samples, n_objectives=self.n_objectives,
            function=self.function, betas=self.betas, length_scales=self.length_scales, batch_size=self.batch_size,
            bounds=self.bounds, callbacks=self.callbacks if self.callbacks else None, acquisition=self.acquisition)

    def pareto_analysis(self) -> np.ndarray:
        """Pareto-efficient objective rows among the evaluated points (:465-488)."""
        evaluated_y = self.y_vector[: self.n_evaluations]
        evaluated_x = self.x_vector[: self.n_evaluations]
        pareto_inputs, pareto_objectives = compute_pareto_front(evaluated_x, evaluated_y)
        print_pareto_analysis(pareto_inputs, pareto_objectives)
        return pareto_objectives
