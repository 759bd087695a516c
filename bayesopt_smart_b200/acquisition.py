"""Drop-in for the reference's ``bayesopt/acquisition.py``: same signatures and in-place semantics,
computed by libbo_b200.so kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import DeviceGP, HviFront, _ptr, _stream, require_cuda, to_device

_F64 = torch.float64
_selector = None


def _get_selector() -> DeviceGP:
    global _selector
    if _selector is None:
        _selector = DeviceGP()
    return _selector


def _ucb_device(mu: torch.Tensor, var: torch.Tensor, betas, want_acq: bool):
    """ucb = mu + beta*sqrt(|var|) on (m, M) device arrays (prior 0 / 1 makes the standardisation exact identity)."""
    lib = _lib.load()
    m, n_cand = mu.shape
    ucb = torch.empty_like(mu)
    acq = torch.empty(n_cand, dtype=_F64, device=mu.device) if want_acq else None
    _, pm = _lib.host_doubles(np.zeros(m), m)
    _, pv = _lib.host_doubles(np.ones(m), m)
    _, pb = _lib.host_doubles(betas, m)
    _lib.check(lib.bo_acquisition_f64(None, None, _ptr(ucb), _ptr(acq), _ptr(mu), _ptr(var), n_cand, n_cand, m, pm,
                                      pv, pb, _stream()))
    return ucb, acq


def upper_confidence_bound(mu: np.ndarray, variance: np.ndarray, beta: float) -> np.ndarray:
    """``mu + beta * sqrt(|variance|)`` for one objective.  Reference acquisition.py:33-52."""
    dev = require_cuda()
    mu_d = to_device(np.asarray(mu, dtype=np.float64).reshape(1, -1), _F64, dev)
    var_d = to_device(np.asarray(variance, dtype=np.float64).reshape(1, -1), _F64, dev)
    ucb, _ = _ucb_device(mu_d, var_d, [beta], False)
    return ucb.cpu().numpy().reshape(np.shape(mu))


def update_ucb(ucb: np.ndarray, mu_objectives: np.ndarray, variance_objectives: np.ndarray, betas: np.ndarray) -> None:
    """UCB per objective, in place.  Reference acquisition.py:55-81."""
    dev = require_cuda()
    out, _ = _ucb_device(to_device(mu_objectives, _F64, dev), to_device(variance_objectives, _F64, dev), betas, False)
    ucb[:, :] = out.cpu().numpy()


def update_hypervolume_improvement(acquisition_values: np.ndarray, ucb: np.ndarray) -> None:
    """``acq[i] = sum_o ucb[o, i]`` (sequential from 0.0), in place.  Reference acquisition.py:89-108."""
    dev = require_cuda()
    u = to_device(ucb, _F64, dev)
    _, acq = _ucb_device(u, torch.zeros_like(u), np.zeros(u.shape[0]), True)
    acquisition_values[:] = acq.cpu().numpy()


def select_next_batch(input_space: np.ndarray, acquisition_values: np.ndarray, evaluated_points: np.ndarray,
                      batch_size: int = 3) -> np.ndarray:
    """Best ``batch_size`` candidates not yet evaluated.  Reference acquisition.py:116-144.

    Ordering: value descending, index ascending on ties (the reference's argsort leaves ties
    unspecified).  Returns ``np.array(batch)`` with ``input_space``'s dtype, possibly fewer rows.
    """
    dev = require_cuda()
    sel = _get_selector()
    cand = to_device(input_space, None, dev)
    acq = to_device(acquisition_values, _F64, dev)
    ev = to_device(np.asarray(evaluated_points, dtype=np.float64).reshape(-1, input_space.shape[1]), _F64, dev)
    _, idx = sel.select(cand, acq, ev, int(batch_size))
    return np.array([input_space[i] for i in idx])


def exact_hvi_device(ucb: torch.Tensor, front: np.ndarray, reference_point: np.ndarray) -> torch.Tensor:
    """Exact hypervolume improvement of each UCB vector against ``front`` (opt-in mode, m = 2 or 3).

    ``HVI(u) = HV(front U {u}) - HV(front)`` with maximisation and ``reference_point`` as the lower
    corner.  The reference has no counterpart (its "HVI" is sum-UCB, acquisition.py:104-108); the
    specification is ``oracle.gp_oracle.exact_hvi``.
    """
    lib = _lib.load()
    m, n_cand = ucb.shape
    f = np.asarray(front, dtype=np.float64).reshape(-1, m)
    f = f[np.argsort(-f[:, 0], kind="stable")]  # objective 0 descending, as the kernel's sweep expects
    f_dev = to_device(f, _F64, ucb.device) if f.shape[0] else None
    out = torch.empty(n_cand, dtype=_F64, device=ucb.device)
    _, pr = _lib.host_doubles(reference_point, m)
    _lib.check(lib.bo_hvi_f64(_ptr(out), _ptr(ucb), ucb.stride(0), n_cand, m, _ptr(f_dev), f.shape[0], pr,
                              _stream()))
    return out


def ucb_and_exact_hvi_device(mu: torch.Tensor, var: torch.Tensor, prior_mean, prior_variance, betas,
                             front: HviFront, want_ucb: bool = True):
    """ONE fused per-candidate pass over (m, M) device arrays: standardise, UCB and the exact hypervolume
    improvement of the UCB vector against a device-prepared front (bo_acquisition_hvi_f64) -- the fused form of
    standardize_objectives + update_ucb + update_hypervolume_improvement (numba_kernels.py:538-570,
    acquisition.py:55-108) with the exact HVI in place of the reference's sum.  Returns (ucb or None, hvi)."""
    lib = _lib.load()
    m, n_cand = mu.shape
    ucb = torch.empty_like(mu) if want_ucb else None
    hvi = torch.empty(n_cand, dtype=_F64, device=mu.device)
    _, pm = _lib.host_doubles(prior_mean, m)
    _, pv = _lib.host_doubles(prior_variance, m)
    _, pb = _lib.host_doubles(betas, m)
    _lib.check(lib.bo_acquisition_hvi_f64(None, None, _ptr(ucb), _ptr(hvi), _ptr(mu), _ptr(var), mu.stride(0), n_cand,
                                          m, pm, pv, pb, _ptr(front.prepared), _ptr(front.count), front.n_points,
                                          front.ref_ptr(), _stream()))
    return ucb, hvi


def update_exact_hypervolume_improvement(acquisition_values: np.ndarray, ucb: np.ndarray, front: np.ndarray,
                                         reference_point: np.ndarray) -> None:
    """In-place host wrapper of :func:`exact_hvi_device`."""
    dev = require_cuda()
    acquisition_values[:] = exact_hvi_device(to_device(ucb, _F64, dev), front, reference_point).cpu().numpy()
