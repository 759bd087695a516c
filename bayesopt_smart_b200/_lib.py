"""ctypes binding of libbo_b200.so (the C ABI declared in include/bo_b200.h).

There is no CPU or Numba fallback: if the shared library is missing, importing any
compute entry point raises with the build instruction.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_uint8, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbo_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "bo_b200.h")

BO_OK, BO_ERR_INVALID, BO_ERR_CUDA, BO_ERR_NOT_PD, BO_ERR_WORKSPACE, BO_ERR_GUARD = 0, 1, 2, 3, 4, 5
BO_CAND_F64, BO_CAND_I64 = 0, 1
BO_MAX_OBJECTIVES, BO_MAX_DIMS, BO_MAX_TOPK, BO_TILE = 4, 16, 1024, 128
BO_MAX_APPEND = 32
BO_PROF_CONTRACTION, BO_PROF_KSTAR, BO_PROF_FINALIZE, BO_PROF_TOPK, BO_PROF_FIT = 0, 1, 2, 3, 4


class BoError(RuntimeError):
    """A libbo_b200 call returned a non-zero status."""


class Int8GuardError(BoError):
    """The INT8 variance engine's sampled cross-check against the FP64 engine exceeded its tolerance."""


_dp = POINTER(c_double)

# name -> (restype, argtypes); mirrors include/bo_b200.h one to one
_SIGNATURES = {
    "bo_abi_version": (c_int, []),
    "bo_last_error": (c_char_p, []),
    "bo_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t)]),
    "bo_gram_f64": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, _dp, _dp, c_void_p]),
    "bo_inverse_workspace_bytes": (c_size_t, [c_int, c_int]),
    "bo_inverse_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_npad": (c_int, [c_int]),
    "bo_last_clamped_pivots": (c_int, []),
    "bo_wpack_doubles": (c_size_t, [c_int]),
    "bo_fit_workspace_bytes": (c_size_t, [c_int, c_int]),
    "bo_gp_fit_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, _dp, _dp,
                              _dp, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_gp_append_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 _dp, _dp, _dp, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_score_workspace_bytes": (c_size_t, [c_int, c_int, c_longlong]),
    "bo_score_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int,
                             c_int, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, _dp, _dp,
                             _dp, _dp, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_i8_wq_bytes": (c_size_t, [c_int]),
    "bo_i8_wscale_doubles": (c_size_t, [c_int, c_int]),
    "bo_i8_quantize_w": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "bo_score_i8_workspace_bytes": (c_size_t, [c_int, c_int, c_longlong]),
    "bo_score_i8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int,
                            c_int, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                            _dp, _dp, _dp, _dp, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_i8_guard_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_longlong, c_longlong]),
    "bo_i8_guard_f64": (c_int, [_dp, _dp, c_void_p, c_int, c_int, c_longlong, c_longlong, c_void_p, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, _dp, _dp, _dp, c_double,
                                c_double, c_double, c_void_p, c_size_t, c_void_p]),
    "bo_i8_peak_tops": (c_int, [_dp, c_double, c_void_p]),
    "bo_i8_kstar_digits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p, c_int, c_int,
                                   c_int, c_int, c_void_p, _dp, _dp, c_void_p]),
    "bo_i8_sumsq": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, _dp,
                            c_void_p]),
    "bo_acquisition_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                   c_longlong, c_int, _dp, _dp, _dp, c_void_p]),
    "bo_topk_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "bo_topk_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_size_t,
                            c_void_p]),
    "bo_match_rows_f64": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_int, c_int, c_void_p, c_int,
                                  c_int, c_int, c_void_p]),
    "bo_mask_evaluated_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p, c_int, c_int,
                                      c_int, c_void_p]),
    "bo_topk_merge_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t,
                                  c_void_p]),
    "bo_pareto_mask_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_int, c_void_p]),
    "bo_pareto_mask_against_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_void_p, c_longlong,
                                           c_longlong, c_int, c_void_p]),
    "bo_pareto_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "bo_pareto_mask_filtered_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_int, c_void_p, c_size_t,
                                            c_void_p]),
    "bo_mll_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bo_mll_batched_f64": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, _dp, _dp, _dp,
                                   c_int, c_void_p, c_size_t, c_void_p]),
    "bo_hvi_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_int, c_void_p, c_int, _dp, c_void_p]),
    "bo_hvi_front_doubles": (c_size_t, [c_int, c_int]),
    "bo_hvi_workspace_bytes": (c_size_t, [c_int, c_int]),
    "bo_hvi_prepare_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, _dp, c_void_p, c_size_t,
                                   c_void_p]),
    "bo_acquisition_hvi_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                       c_longlong, c_int, _dp, _dp, _dp, c_void_p, c_void_p, c_int, _dp, c_void_p]),
    "bo_score_hvi_f64": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                 c_void_p, c_int, c_int, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, _dp, _dp, _dp, _dp, c_double, c_void_p, c_void_p, c_int, _dp,
                                 c_void_p, c_size_t, c_void_p]),
    "bo_grid_i64": (c_int, [c_void_p, c_longlong, POINTER(c_longlong), POINTER(c_longlong), c_int, c_longlong,
                            c_longlong, c_void_p]),
    "bo_kstar_dense_f64": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_int, c_void_p, c_int, c_int,
                                   c_longlong, c_int, c_int, c_int, c_int, _dp, _dp, c_void_p]),
    "bo_dense_workspace_bytes": (c_size_t, [c_int, c_longlong]),
    "bo_mean_dense_f64": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_longlong, c_void_p, c_int,
                                  c_longlong, c_void_p, c_int, _dp, c_int, c_longlong, c_int, c_void_p, c_size_t,
                                  c_void_p]),
    "bo_variance_dense_f64": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_longlong, c_void_p, c_int,
                                      c_longlong, _dp, c_double, c_int, c_longlong, c_int, c_void_p, c_size_t,
                                      c_void_p]),
    "bo_dgemm_nt_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "bo_launch_count": (c_longlong, [c_int]),
    "bo_profile_enable": (c_int, [c_int]),
    "bo_profile_read": (c_int, [_dp, POINTER(c_longlong), _dp]),
    "bo_profile_read_kernel": (c_int, [c_int, _dp, POINTER(c_longlong), _dp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> ctypes.CDLL:
    """Load libbo_b200.so (once) and attach the prototypes.  Fails loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BoError(
            f"{LIB_PATH} is missing: the CUDA library is not built. Run `python -m bayesopt_smart_b200.build` "
            "(nvcc, sm_100a). bayesopt_smart_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if lib.bo_abi_version() != 1:
        raise BoError(f"ABI version mismatch: library {lib.bo_abi_version()}, binding 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == BO_OK:
        return
    msg = load().bo_last_error().decode("utf-8", "replace")
    if rc == BO_ERR_NOT_PD:
        raise np.linalg.LinAlgError(msg or "Matrix is not positive definite")
    if rc == BO_ERR_GUARD:
        raise Int8GuardError(msg)
    raise BoError(f"libbo_b200 error {rc}: {msg}")


def host_doubles(values, m: int):
    """Small per-objective host array -> (keep-alive ndarray, double*)."""
    a = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
    if a.size < m:
        raise ValueError(f"expected at least {m} values, got {a.size}")
    return a, a.ctypes.data_as(_dp)
