"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol the
header declares, argument validation works without a GPU, and the Python surface mirrors the
reference's signatures.  No compute call is made here."""
import inspect
import os
import re

import numpy as np
import pytest

import bayesopt_smart_b200 as pkg
from bayesopt_smart_b200 import _lib, acquisition, bayesian_optimization, numba_kernels, pareto


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from bayesopt_smart_b200.build import build

        build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    text = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(bo_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_geometry(lib):
    assert lib.bo_abi_version() == 1
    assert lib.bo_npad(1) == 128 and lib.bo_npad(128) == 128 and lib.bo_npad(129) == 256
    # packed W: 8 k-tiles per (row block, k block) pair on or below the diagonal, 2048 doubles each
    assert lib.bo_wpack_doubles(128) == 8 * 2048
    assert lib.bo_wpack_doubles(1024) == 8 * (8 * 9 // 2) * 2048
    assert lib.bo_fit_workspace_bytes(1024, 2) > 2 * 2 * 1024 * 1024 * 8
    assert lib.bo_score_workspace_bytes(1024, 2, 10**6) > 0
    assert lib.bo_topk_workspace_bytes(10**6, 19) > 0
    assert lib.bo_mll_workspace_bytes(256, 2, 4) > 0


def test_argument_validation_without_gpu(lib):
    rc = lib.bo_topk_f64(None, None, None, 10, 3, 0, None, 0, None)
    assert rc == _lib.BO_ERR_INVALID
    assert b"null pointer" in lib.bo_last_error()
    with pytest.raises(_lib.BoError):
        _lib.check(rc)
    a, p = _lib.host_doubles([1.0, 2.0], 2)
    rc = lib.bo_hvi_f64(1, 1, 1, 1, 4, 1, 0, p, None)
    assert rc == _lib.BO_ERR_INVALID


def test_not_pd_maps_to_linalgerror(lib):
    with pytest.raises(np.linalg.LinAlgError):
        _lib.check(_lib.BO_ERR_NOT_PD)


REFERENCE_SIGNATURES = {
    (numba_kernels, "initialize_lhs_integer"): ["x_vector", "y_vector", "bounds", "function", "n_samples"],
    (numba_kernels, "compute_prior_mean"): ["y_vector", "n_evaluations", "n_objectives"],
    (numba_kernels, "compute_prior_variance"): ["y_vector", "n_evaluations", "n_objectives"],
    (numba_kernels, "compute_mll"): ["x_vector", "y_vector", "kernel_matrix", "prior_mean", "prior_variance",
                                     "length_scales", "current_eval"],
    (numba_kernels, "optimize_hyperparams_mll"): ["x_vector", "y_vector", "kernel_matrix", "prior_mean",
                                                  "prior_variance", "length_scales", "current_eval"],
    (numba_kernels, "update_k"): ["kernel_matrix", "x_vector", "last_eval", "current_eval", "prior_variance",
                                  "length_scales"],
    (numba_kernels, "invert_k"): ["current_eval", "kernel_matrix"],
    (numba_kernels, "update_k_star"): ["k_star", "x_vector", "input_space", "last_eval", "current_eval",
                                       "prior_variance", "length_scales"],
    (numba_kernels, "update_mean"): ["mu_objectives", "k_star", "inverted_kernel_matrix", "y_vector", "prior_mean",
                                     "current_eval"],
    (numba_kernels, "update_variance"): ["variance_objectives", "k_star", "inverted_kernel_matrix", "prior_variance",
                                         "current_eval"],
    (numba_kernels, "standardize_objectives"): ["std_mu_objectives", "std_variance_objectives", "mu_objectives",
                                                "variance_objectives", "prior_mean", "prior_variance"],
    (acquisition, "upper_confidence_bound"): ["mu", "variance", "beta"],
    (acquisition, "update_ucb"): ["ucb", "mu_objectives", "variance_objectives", "betas"],
    (acquisition, "update_hypervolume_improvement"): ["acquisition_values", "ucb"],
    (acquisition, "select_next_batch"): ["input_space", "acquisition_values", "evaluated_points", "batch_size"],
    (pareto, "is_pareto_efficient"): ["y_vector"],
    (pareto, "compute_pareto_front"): ["x_vector", "y_vector"],
    (pareto, "print_pareto_analysis"): ["pareto_inputs", "pareto_objectives"],
}


@pytest.mark.parametrize("key", list(REFERENCE_SIGNATURES), ids=lambda k: k[1])
def test_python_surface_matches_reference(key):
    """Parameter names and order of the reference's free functions."""
    mod, name = key
    params = list(inspect.signature(getattr(mod, name)).parameters)
    assert params == REFERENCE_SIGNATURES[key]


def test_optimize_signature_and_exports():
    params = list(inspect.signature(bayesian_optimization.optimize).parameters)
    ref = ["x_vector", "y_vector", "kernel_matrices", "k_star", "mu_objectives", "variance_objectives",
           "std_mu_objectives", "std_variance_objectives", "ucb", "acquisition_values", "input_space", "prior_mean",
           "prior_variance", "reference_point", "n_evaluations", "total_samples", "n_objectives", "function", "betas",
           "length_scales", "batch_size", "bounds", "callbacks"]
    assert params[: len(ref)] == ref  # bayesian_optimization.py:51-75; one trailing opt-in kwarg is allowed
    ctor = list(inspect.signature(pkg.BayesianOptimization.__init__).parameters)
    assert ctor == ["self", "function", "bounds", "n_objectives", "n_iterations", "kwargs"]
    for name in ["BayesianOptimization", "PlotterCallback", "ProgressLogger", "OptimizationLogger",
                 "PerformanceMonitor", "select_next_batch", "is_pareto_efficient", "compute_pareto_front",
                 "print_pareto_analysis"]:
        assert hasattr(pkg, name)
    from bayesopt_smart_b200 import config

    assert (config.KERNEL_JITTER, config.CHOLESKY_JITTER, config.MIN_VARIANCE) == (1e-6, 1e-8, 1e-10)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BoError):
        pkg.is_pareto_efficient(np.zeros((4, 2)))
    with pytest.raises(_lib.BoError):
        pkg.DeviceGP()


def test_device_grid_is_used_only_for_the_untouched_integer_grid():
    """the device-generated candidate grid replaces the upload only when input_space is provably the
    int64 grid of integral bounds (no GPU needed: the constructor is host-only)."""
    f = lambda p: np.array([p[0], -p[1]])  # noqa: E731
    bo = pkg.BayesianOptimization(f, [(0, 7), (2, 9)], n_objectives=2, n_iterations=1, initial_samples=3)
    assert bo._input_space_is_the_integer_grid()
    bo.input_space[5, 1] += 1  # edited in place: same object, different points
    assert not bo._input_space_is_the_integer_grid()
    bo.input_space[5, 1] -= 1
    bo.input_space = bo.input_space.copy()  # replaced by the caller
    assert not bo._input_space_is_the_integer_grid()
    bo2 = pkg.BayesianOptimization(f, [(0.5, 4.5), (0, 3)], n_objectives=2, n_iterations=1, initial_samples=3)
    assert bo2.input_space.dtype == np.float64  # np.arange over non-integral bounds gives a float grid
    assert not bo2._input_space_is_the_integer_grid()
