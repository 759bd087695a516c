"""Drop-in for the reference's ``bayesopt/pareto.py``; the O(n^2) dominance test runs on the GPU."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream, require_cuda, to_device

_F64 = torch.float64
_DIRECT_LIMIT = 1 << 16  # below this the plain n x n kernel is used


def _mask_direct(y: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n, m = y.shape
    mask = torch.empty(n, dtype=torch.uint8, device=y.device)
    if n:
        _lib.check(lib.bo_pareto_mask_f64(_ptr(mask), _ptr(y), y.stride(0), n, m, _stream()))
    return mask


def _mask_against(y: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n, m = y.shape
    mask = torch.empty(n, dtype=torch.uint8, device=y.device)
    if n:
        _lib.check(lib.bo_pareto_mask_against_f64(_ptr(mask), _ptr(y), y.stride(0), n, _ptr(z), z.stride(0),
                                                  z.shape[0], m, _stream()))
    return mask


_ws = {}


def pareto_mask_device(y: torch.Tensor) -> torch.Tensor:
    """Non-dominated mask (uint8) of an (n, m) CUDA tensor, maximisation, duplicates / NaN rows kept.

    Up to 65 536 rows: the plain n x n warp-ballot kernel.  Larger sets (BASELINE config 4 filters 8 M UCB
    vectors): ``bo_pareto_mask_filtered_f64`` -- sampling, sample-front ordering, thinning, stream compaction
    and the final test all run inside the library on the device, with no host synchronisation and no PyTorch
    sort / boolean indexing.  The result equals the direct n x n test.
    """
    y = y.contiguous()
    n, m = y.shape
    if n <= _DIRECT_LIMIT:
        return _mask_direct(y)
    lib = _lib.load()
    ws_bytes = lib.bo_pareto_workspace_bytes(n, m)
    ws = _ws.get(y.device)
    if ws is None or ws.numel() < ws_bytes:
        ws = _ws[y.device] = torch.empty(int(ws_bytes), dtype=torch.uint8, device=y.device)
    mask = torch.empty(n, dtype=torch.uint8, device=y.device)
    _lib.check(lib.bo_pareto_mask_filtered_f64(_ptr(mask), _ptr(y), y.stride(0), n, m, _ptr(ws), ws_bytes, _stream()))
    return mask


def is_pareto_efficient(y_vector: np.ndarray) -> np.ndarray:
    """Boolean mask of Pareto-efficient rows (maximisation).  Reference pareto.py:12-45."""
    dev = require_cuda()
    y = np.asarray(y_vector, dtype=np.float64)
    if y.ndim != 2:
        raise ValueError("y_vector must be (n_points, n_objectives)")
    if y.shape[1] > _lib.BO_MAX_OBJECTIVES:
        raise ValueError(f"at most {_lib.BO_MAX_OBJECTIVES} objectives")
    if y.shape[0] == 0:
        return np.ones(0, dtype=bool)
    return pareto_mask_device(to_device(y, _F64, dev)).cpu().numpy().astype(bool)


def compute_pareto_front(x_vector: np.ndarray, y_vector: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Rows of (x, y) on the Pareto front, input order preserved.  Reference pareto.py:48-64."""
    is_efficient = is_pareto_efficient(y_vector)
    return x_vector[is_efficient], y_vector[is_efficient]


def print_pareto_analysis(pareto_inputs: np.ndarray, pareto_objectives: np.ndarray) -> None:
    """Reference pareto.py:67-80."""
    print("Pareto Analysis Results:")
    for i, (input_point, obj_values) in enumerate(zip(pareto_inputs, pareto_objectives)):
        print(f"Input: {input_point}, Pareto Point {i + 1}: {obj_values}")
