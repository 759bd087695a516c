"""Device-resident GP scorer: the fused form of the reference's per-iteration hot path.

``DeviceGP.fit``    = update_k + invert_k (+ the Kinv @ dy half of update_mean)
                      (reference bayesian_optimization.py:129-142)
``DeviceGP.score``  = update_k_star + update_mean + update_variance + standardize_objectives
                      + update_ucb + update_hypervolume_improvement (:145-199)
``DeviceGP.select`` = select_next_batch (:202-207)

PyTorch is used only for plumbing (device memory, pinned staging buffers, streams); every
arithmetic step is a kernel of libbo_b200.so reached through the C ABI.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .config import KERNEL_JITTER, MIN_VARIANCE

_F64 = torch.float64


def require_cuda() -> torch.device:
    """The package is GPU-only: fail loudly instead of falling back to the CPU."""
    if not torch.cuda.is_available():
        raise _lib.BoError("bayesopt_smart_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Workspace:
    """Grow-only device scratch buffers, one per purpose."""

    def __init__(self):
        self._bufs: Dict[str, torch.Tensor] = {}

    def get(self, key: str, nbytes: int, device) -> torch.Tensor:
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != device:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


def to_device(a, dtype=None, device=None) -> torch.Tensor:
    """numpy / torch -> contiguous CUDA tensor.  Host tensors that are already pinned copy asynchronously;
    NumPy arrays go through the driver's own staging (pinning per call would cost more than it saves)."""
    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=dtype or a.dtype)
        return t.contiguous()
    arr = np.ascontiguousarray(a)
    if dtype is not None:
        arr = arr.astype({torch.float64: np.float64, torch.int64: np.int64}[dtype], copy=False)
    return torch.from_numpy(arr).to(device)


def grid_candidates(bounds, row0: int = 0, rows: Optional[int] = None, device=None) -> torch.Tensor:
    """Rows [row0, row0 + rows) of the reference's candidate set -- every integer point of the box ``bounds``
    (upper bounds exclusive), C order (bayesian_optimization.py:338-340) -- generated on the device."""
    device = device or require_cuda()
    lo = np.ascontiguousarray([int(b[0]) for b in bounds], dtype=np.int64)
    hi = np.ascontiguousarray([int(b[1]) for b in bounds], dtype=np.int64)
    total = int(np.prod(hi - lo))
    rows = total - row0 if rows is None else int(rows)
    if np.any(hi <= lo) or row0 < 0 or rows < 0 or row0 + rows > total:
        raise _lib.BoError(f"grid_candidates: bad bounds / row range ({bounds}, row0={row0}, rows={rows})")
    out = torch.empty((rows, len(lo)), dtype=torch.int64, device=device)
    if rows == 0:
        return out
    ll = ctypes.POINTER(ctypes.c_longlong)
    _lib.check(_lib.load().bo_grid_i64(_ptr(out), out.stride(0), lo.ctypes.data_as(ll),
                                       hi.ctypes.data_as(ll), len(lo), int(row0), rows, _stream()))
    return out


def candidate_kind(cand: torch.Tensor) -> int:
    if cand.dtype == torch.float64:
        return _lib.BO_CAND_F64
    if cand.dtype == torch.int64:
        return _lib.BO_CAND_I64
    raise TypeError(f"candidates must be float64 or int64, got {cand.dtype}")


VARIANCE_ENGINES = ("dmma", "int8")


class HviFront:
    """A Pareto front prepared ON THE DEVICE for the fused UCB + exact-HVI pass (bo_hvi_prepare_f64): points
    clipped to the reference point, dominated points removed by the dominance kernel, sorted by objective 0
    descending, plus the staircase prefix areas (m = 2) or objective-2 levels (m = 3).  No host sort, no host
    synchronisation; ``count`` stays a device scalar."""

    def __init__(self, points, reference_point, device=None):
        device = device or require_cuda()
        lib = _lib.load()
        pts = to_device(np.asarray(points, dtype=np.float64) if not isinstance(points, torch.Tensor) else points,
                        _F64, device)
        if pts.dim() != 2 or pts.shape[1] not in (2, 3):
            raise ValueError("exact HVI needs (n, 2) or (n, 3) points")
        self.m = int(pts.shape[1])
        self.n_points = int(pts.shape[0])
        self.ref, pr = _lib.host_doubles(reference_point, self.m)
        self.prepared = torch.empty(lib.bo_hvi_front_doubles(self.n_points, self.m), dtype=_F64, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        ws_bytes = lib.bo_hvi_workspace_bytes(self.n_points, self.m)
        ws = torch.empty(max(int(ws_bytes), 256), dtype=torch.uint8, device=device)
        _lib.check(lib.bo_hvi_prepare_f64(_ptr(self.prepared), _ptr(self.count), _ptr(pts) if self.n_points else None,
                                          pts.stride(0) if self.n_points else self.m, self.n_points, self.m, pr,
                                          _ptr(ws), ws_bytes, _stream()))
        self._keep = (pts, ws)  # alive until the stream has consumed them

    def ref_ptr(self):
        return self.ref.ctypes.data_as(_lib._dp)

    def points(self) -> np.ndarray:
        """The prepared front as a host array (P, m), objective 0 descending (diagnostics / tests)."""
        p = int(self.count.item())
        cap = max(self.n_points, 1)
        cols = [self.prepared[o * cap: o * cap + p].cpu().numpy() for o in range(self.m)]
        return np.stack(cols, axis=1) if p else np.zeros((0, self.m))


class DeviceGP:
    """GP factor + scorer living in HBM.  One instance per process / GPU.

    ``variance_engine`` picks how |W k*|^2 is contracted: ``"dmma"`` (FP64 tensor cores, the default) or
    ``"int8"`` (error-free digit splitting on tcgen05 INT8 tensor cores, ~1e-12 of the prior variance away
    from the FP64 result).  ``BO_VARIANCE_ENGINE`` overrides the default.
    """

    def __init__(self, device=None, variance_engine: Optional[str] = None, int8_guard_tol: Optional[float] = None,
                 int8_guard_stride: int = 4096):
        self.device = device or require_cuda()
        self.variance_engine = variance_engine or os.environ.get("BO_VARIANCE_ENGINE", "dmma")
        if self.variance_engine not in VARIANCE_ENGINES:
            raise ValueError(f"variance_engine must be one of {VARIANCE_ENGINES}, got {self.variance_engine!r}")
        # INT8 engine: every score() cross-checks one candidate in `int8_guard_stride` against the FP64 engine and
        # raises Int8GuardError -- no fallback -- if |d var| / prior_variance exceeds the tolerance (default: the 1e-9
        # parity target).  BO_I8_GUARD=0 or int8_guard_tol <= 0 switches the check off.
        if int8_guard_tol is None:
            int8_guard_tol = 0.0 if os.environ.get("BO_I8_GUARD", "1") == "0" else 1e-9
        self.int8_guard_tol = float(int8_guard_tol)
        self.int8_guard_stride = max(1, int(int8_guard_stride))
        self.last_guard_worst: Optional[float] = None  # largest sampled |d var| / prior_variance of the last score()
        self.last_guard_tolerance: Optional[float] = None  # what it was held to: max(tol, 10 eps cond_upper)
        self.wq: Optional[torch.Tensor] = None
        self.wscale: Optional[torch.Tensor] = None
        self.lib = _lib.load()
        self.ws = _Workspace()
        self.n = 0
        self.d = 0
        self.m = 0
        self.x: Optional[torch.Tensor] = None
        self.wpack: Optional[torch.Tensor] = None
        self.alpha: Optional[torch.Tensor] = None
        self.prior_mean = self.prior_variance = self.length_scales = None
        self.clamped_pivots = 0  # Cholesky pivots the last fit clamped to the jitter (0 on well-conditioned input)
        self.last_fit = None     # "full" or "append": how the last fit() produced the factor
        self.y: Optional[torch.Tensor] = None
        self._fit_key = None     # (prior_mean, prior_variance, length_scales, jitter) of the resident factor
        # resident training rows: device buffers with spare capacity + the host copy they were uploaded from
        self._x_res: Optional[torch.Tensor] = None
        self._y_res: Optional[torch.Tensor] = None
        self._x_host: Optional[np.ndarray] = None
        self._y_host: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ fit
    def _stage_training(self, x_vector, y_vector, n: int, compare: bool = True):
        """Training rows -> device tensors (n, d), (n, m), plus whether the rows of the resident factor are an
        unchanged prefix of them.  Host (NumPy) inputs are compared with the host copy kept from the last fit and
        only the NEW rows are uploaded into the resident buffers; device tensors are compared on the device."""
        n_old = self.n if self.wpack is not None else 0
        if isinstance(x_vector, torch.Tensor) or isinstance(y_vector, torch.Tensor):
            x = to_device(x_vector, _F64, self.device)
            y = to_device(y_vector, _F64, self.device)
            if x.dim() != 2 or y.dim() != 2 or x.shape[0] < n or y.shape[0] < n:
                raise ValueError("x_vector (T,d) and y_vector (T,m) must hold at least current_eval rows")
            same = bool(compare and 0 < n_old <= n and self.x is not None and x.shape[1] == self.x.shape[1]
                        and y.shape[1] == self.y.shape[1] and torch.equal(x[:n_old], self.x[:n_old])
                        and torch.equal(y[:n_old], self.y[:n_old]))
            self._x_host = self._y_host = None
            return x[:n].clone(), y[:n].clone(), same
        xh = np.ascontiguousarray(np.asarray(x_vector, dtype=np.float64))
        yh = np.ascontiguousarray(np.asarray(y_vector, dtype=np.float64))
        if xh.ndim != 2 or yh.ndim != 2 or xh.shape[0] < n or yh.shape[0] < n:
            raise ValueError("x_vector (T,d) and y_vector (T,m) must hold at least current_eval rows")
        xh, yh = xh[:n], yh[:n]
        hx, hy = self._x_host, self._y_host
        same = bool(0 < n_old <= n and hx is not None and hx.shape[1] == xh.shape[1] and hy.shape[1] == yh.shape[1]
                    and np.array_equal(xh[:n_old], hx[:n_old]) and np.array_equal(yh[:n_old], hy[:n_old]))
        res_x, res_y = self._x_res, self._y_res
        if same and res_x is not None and res_x.shape[0] >= n:
            if n > n_old:  # only the new rows cross PCIe; the prefix is already resident
                res_x[n_old:n].copy_(torch.from_numpy(xh[n_old:n]))
                res_y[n_old:n].copy_(torch.from_numpy(yh[n_old:n]))
        else:
            cap = self.lib.bo_npad(n) + _lib.BO_TILE
            res_x = torch.empty((cap, xh.shape[1]), dtype=_F64, device=self.device)
            res_y = torch.empty((cap, yh.shape[1]), dtype=_F64, device=self.device)
            res_x[:n].copy_(torch.from_numpy(xh))
            res_y[:n].copy_(torch.from_numpy(yh))
            self._x_res, self._y_res = res_x, res_y
        self._x_host, self._y_host = xh.copy(), yh.copy()
        return res_x[:n], res_y[:n], same

    def _can_append(self, n: int, d: int, m: int, hyper_key) -> bool:
        """The resident factor can be extended by the rows [self.n, n): same hyper-parameters bit
        for bit, same 128-row padding, at most BO_MAX_APPEND new rows, and no clamped pivots in the factor."""
        if self.wpack is None or self._fit_key is None:
            return False
        n_old = self.n
        if not (0 < n_old < n <= n_old + _lib.BO_MAX_APPEND) or self.lib.bo_npad(n) != self.lib.bo_npad(n_old):
            return False
        if self.clamped_pivots or d != self.d or m != self.m:
            return False  # a factor that needed pivot clamping (cond ~ 1e15) is not extended, it is rebuilt
        return hyper_key == self._fit_key

    def fit(self, x_vector, y_vector, prior_mean, prior_variance, length_scales, current_eval: int,
            jitter: float = KERNEL_JITTER, incremental: bool = True) -> None:
        """Factor K + jitter I for the first ``current_eval`` rows.  Raises numpy LinAlgError if not PD.

        The training rows, the dense factor and ``W = L^-1`` stay resident in HBM between calls.  ``incremental``
        (default on): when the previous fit of this object used bit-identical hyper-parameters and its training
        rows are an unchanged prefix of the new ones, the factor is EXTENDED by the new rows
        (``bo_gp_append_f64``, O(b n^2)) instead of rebuilt (O(n^3)); ``self.last_fit`` says which happened
        ("full" or "append").  The reference always rebuilds (bayesian_optimization.py:129-142); the extended
        factor equals the rebuilt one up to rounding (tests/test_gpu_append.py).
        """
        n = int(current_eval)
        d, m = int(np.shape(x_vector)[1]), int(np.shape(y_vector)[1])
        self._mean_h, pm = _lib.host_doubles(prior_mean, m)
        self._var_h, pv = _lib.host_doubles(prior_variance, m)
        self._ls_h, pl = _lib.host_doubles(length_scales, m)
        hyper_key = (tuple(self._mean_h[:m].tolist()), tuple(self._var_h[:m].tolist()), tuple(self._ls_h[:m].tolist()),
                     float(jitter))
        may_append = bool(incremental and self._can_append(n, d, m, hyper_key))
        # the prefix comparison of device tensors costs a host synchronisation: only done when an append is possible
        x, y, prefix_same = self._stage_training(x_vector, y_vector, n, compare=may_append)
        npad = self.lib.bo_npad(n)
        ws_bytes = self.lib.bo_fit_workspace_bytes(n, m)
        appended = False
        if may_append and prefix_same:
            ws = self.ws.get("fit", ws_bytes, self.device)  # same size as at the last fit: the buffer (L, W) is reused
            try:
                _lib.check(self.lib.bo_gp_append_f64(_ptr(self.wpack), _ptr(self.alpha), _ptr(x), x.stride(0), _ptr(y),
                                                     y.stride(0), self.n, n, d, m, pm, pv, pl, float(jitter), _ptr(ws),
                                                     ws_bytes, _stream()))
                appended = True
            except np.linalg.LinAlgError:
                appended = False  # the Schur complement lost definiteness in rounding: rebuild from scratch
        if not appended:
            self.wpack = torch.empty(m * self.lib.bo_wpack_doubles(n), dtype=_F64, device=self.device)
            self.alpha = torch.empty(m * npad, dtype=_F64, device=self.device)
            ws = self.ws.get("fit", ws_bytes, self.device)
            _lib.check(self.lib.bo_gp_fit_f64(_ptr(self.wpack), _ptr(self.alpha), _ptr(x), x.stride(0), _ptr(y),
                                              y.stride(0), n, d, m, pm, pv, pl, float(jitter), _ptr(ws), ws_bytes,
                                              _stream()))
        self.last_fit = "append" if appended else "full"
        self.clamped_pivots = int(self.lib.bo_last_clamped_pivots())
        if self.variance_engine == "int8":
            self.wq = torch.empty(m * self.lib.bo_i8_wq_bytes(n), dtype=torch.uint8, device=self.device)
            self.wscale = torch.empty(self.lib.bo_i8_wscale_doubles(n, m), dtype=_F64, device=self.device)
            _lib.check(self.lib.bo_i8_quantize_w(_ptr(self.wq), _ptr(self.wscale), _ptr(self.wpack), n, m,
                                                 _stream()))
        self.x, self.y = x, y
        self.n, self.d, self.m = n, d, m
        self.prior_mean = self._mean_h.copy()
        self.prior_variance = self._var_h.copy()
        self.length_scales = self._ls_h.copy()
        self._fit_key = hyper_key

    # ------------------------------------------------------------------ score
    def score(self, candidates, betas, *, want=("mu", "var", "acq"), out: Optional[Dict[str, torch.Tensor]] = None,
              min_variance: float = MIN_VARIANCE, hvi: Optional["HviFront"] = None,
              guard: bool = True) -> Dict[str, torch.Tensor]:
        """Posterior + UCB + acquisition for every candidate row.  Returns CUDA tensors.

        ``want`` picks which arrays are written: mu, var, std_mu, std_var, ucb (each (m, M)), acq (M,).
        ``out`` may carry preallocated tensors under the same keys.  ``acq`` is the reference's sum-UCB
        (acquisition.py:104-108) unless ``hvi`` carries a prepared front: then the epilogue of the same pass
        writes the exact hypervolume improvement of each UCB vector (opt-in mode, m = 2 or 3).
        """
        if self.wpack is None:
            raise _lib.BoError("DeviceGP.score called before fit")
        cand = to_device(candidates, None, self.device)
        if cand.dim() != 2 or cand.shape[1] != self.d:
            raise ValueError(f"candidates must be (M, {self.d})")
        kind = candidate_kind(cand)
        n_cand = cand.shape[0]
        m = self.m
        if hvi is not None and hvi.m != m:
            raise ValueError(f"the prepared front has {hvi.m} objectives, the model {m}")
        res: Dict[str, Optional[torch.Tensor]] = {}
        for key in ("mu", "var", "std_mu", "std_var", "ucb", "acq"):
            if out is not None and key in out:
                res[key] = out[key]
            elif key in want:
                shape = (n_cand,) if key == "acq" else (m, n_cand)
                res[key] = torch.empty(shape, dtype=_F64, device=self.device)
            else:
                res[key] = None
        _, pm = _lib.host_doubles(self.prior_mean, m)
        _, pv = _lib.host_doubles(self.prior_variance, m)
        _, pl = _lib.host_doubles(self.length_scales, m)
        bet, pb = _lib.host_doubles(betas, m)
        int8 = self.variance_engine == "int8"
        if int8:
            ws_bytes = self.lib.bo_score_i8_workspace_bytes(self.n, m, n_cand)
            ws = self.ws.get("score_i8", ws_bytes, self.device)
        else:
            ws_bytes = self.lib.bo_score_workspace_bytes(self.n, m, n_cand)
            ws = self.ws.get("score", ws_bytes, self.device)
        # (m, M) outputs may be column slices of wider arrays (slice-wise scoring with overlapped copies): the
        # leading dimension is their common row stride
        ld_out = n_cand
        for key in ("mu", "var", "std_mu", "std_var", "ucb"):
            t = res[key]
            if t is None:
                continue
            if tuple(t.shape) != (m, n_cand) or (n_cand > 1 and t.stride(1) != 1):
                raise ValueError(f"output {key!r} must be ({m}, {n_cand}) with unit column stride")
            ld = t.stride(0) if m > 1 else max(t.stride(0), n_cand)
            if ld_out not in (n_cand, ld) or ld < n_cand:
                raise ValueError("all (m, M) outputs must share one leading dimension >= M")
            ld_out = ld
        if res["acq"] is not None and (res["acq"].numel() != n_cand or (n_cand > 1 and res["acq"].stride(0) != 1)):
            raise ValueError(f"output 'acq' must be a contiguous vector of {n_cand} values")
        outs = (_ptr(res["mu"]), _ptr(res["var"]), _ptr(res["std_mu"]), _ptr(res["std_var"]), _ptr(res["ucb"]),
                _ptr(res["acq"]), ld_out, _ptr(cand), kind, cand.stride(0), n_cand, _ptr(self.x), self.x.stride(0),
                self.n, self.d, m)
        if hvi is not None:
            _lib.check(self.lib.bo_score_hvi_f64(1 if int8 else 0, *outs, _ptr(self.wq) if int8 else _ptr(self.wpack),
                                                 _ptr(self.wscale) if int8 else None, _ptr(self.alpha), pm, pv, pl, pb,
                                                 float(min_variance), _ptr(hvi.prepared), _ptr(hvi.count),
                                                 hvi.n_points, hvi.ref_ptr(), _ptr(ws), ws_bytes, _stream()))
        elif int8:
            _lib.check(self.lib.bo_score_i8(*outs, _ptr(self.wq), _ptr(self.wscale), _ptr(self.alpha), pm, pv, pl, pb,
                                            float(min_variance), _ptr(ws), ws_bytes, _stream()))
        else:
            _lib.check(self.lib.bo_score_f64(*outs, _ptr(self.wpack), _ptr(self.alpha), pm, pv, pl, pb,
                                             float(min_variance), _ptr(ws), ws_bytes, _stream()))
        if guard:
            self.check_int8_guard(cand, min_variance)
        return {k: v for k, v in res.items() if v is not None}

    def check_int8_guard(self, candidates: torch.Tensor, min_variance: float = MIN_VARIANCE) -> None:
        """INT8 engine only (no-op otherwise): score one candidate per ``int8_guard_stride`` with BOTH engines from
        the resident factor and raise ``Int8GuardError`` if they differ by more than the guard tolerance.  ``score``
        calls this itself unless told ``guard=False`` (slice-wise callers check the whole set once)."""
        n_cand = candidates.shape[0]
        if self.variance_engine != "int8" or self.int8_guard_tol <= 0.0 or n_cand == 0:
            return
        m = self.m
        _, pm = _lib.host_doubles(self.prior_mean, m)
        _, pv = _lib.host_doubles(self.prior_variance, m)
        _, pl = _lib.host_doubles(self.length_scales, m)
        gbytes = self.lib.bo_i8_guard_workspace_bytes(self.n, m, self.d, n_cand, self.int8_guard_stride)
        gws = self.ws.get("i8_guard", gbytes, self.device)
        worst, tau = ctypes.c_double(0.0), ctypes.c_double(0.0)
        rc = self.lib.bo_i8_guard_f64(ctypes.byref(worst), ctypes.byref(tau), _ptr(candidates),
                                      candidate_kind(candidates), candidates.stride(0), n_cand, self.int8_guard_stride,
                                      _ptr(self.x), self.x.stride(0), self.n, self.d, m, _ptr(self.wq),
                                      _ptr(self.wscale), _ptr(self.wpack), _ptr(self.alpha), pm, pv, pl,
                                      float(self._fit_key[3]), float(min_variance), self.int8_guard_tol, _ptr(gws),
                                      gbytes, _stream())
        self.last_guard_worst, self.last_guard_tolerance = worst.value, tau.value
        _lib.check(rc)  # Int8GuardError: the caller decides (there is no silent switch to the FP64 engine)

    # ------------------------------------------------------------------ select
    def topk(self, acq: torch.Tensor, k: int, index_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top-k (value desc, index asc, NaN last) of a device vector -> (values, indices) on device."""
        n = acq.numel()
        k = int(min(k, n, _lib.BO_MAX_TOPK))
        vals = torch.empty(k, dtype=_F64, device=self.device)
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        ws_bytes = self.lib.bo_topk_workspace_bytes(n, k)
        ws = self.ws.get("topk", ws_bytes, self.device)
        _lib.check(self.lib.bo_topk_f64(_ptr(vals), _ptr(idx), _ptr(acq), n, k, int(index_base), _ptr(ws), ws_bytes,
                                        _stream()))
        return vals, idx

    def match_rows(self, idx: torch.Tensor, candidates: torch.Tensor, evaluated: torch.Tensor,
                   index_base: int = 0) -> torch.Tensor:
        """flag[i] = candidates[idx[i] - index_base] equals some row of ``evaluated`` (acquisition.py:139)."""
        flags = torch.zeros(idx.numel(), dtype=torch.uint8, device=self.device)
        n_ev = evaluated.shape[0]
        _lib.check(self.lib.bo_match_rows_f64(_ptr(flags), _ptr(idx), idx.numel(), int(index_base), _ptr(candidates),
                                              candidate_kind(candidates), candidates.stride(0),
                                              _ptr(evaluated) if n_ev else None,
                                              evaluated.stride(0) if n_ev else 1, n_ev, candidates.shape[1],
                                              _stream()))
        return flags

    def mask_evaluated(self, acq: torch.Tensor, candidates: torch.Tensor, evaluated: torch.Tensor) -> torch.Tensor:
        """Copy of ``acq`` with NaN (ranked last) at every candidate row that equals an evaluated row -- the
        exhaustive form of the exclusion test of acquisition.py:139 (O(M * n) compares; fallback only)."""
        out = torch.empty_like(acq)
        n_ev = evaluated.shape[0]
        _lib.check(self.lib.bo_mask_evaluated_f64(_ptr(out), _ptr(acq), _ptr(candidates), candidate_kind(candidates),
                                                  candidates.stride(0), acq.numel(),
                                                  _ptr(evaluated) if n_ev else None,
                                                  evaluated.stride(0) if n_ev else 1, n_ev, candidates.shape[1],
                                                  _stream()))
        return out

    def select_listed(self, candidates: torch.Tensor, acq: torch.Tensor, evaluated: torch.Tensor, k: int,
                      index_base: int = 0):
        """Top-``k`` list of this shard with evaluated rows marked: (values, indices, flags) on the device."""
        vals, idx = self.topk(acq, k, index_base)
        flags = self.match_rows(idx, candidates, evaluated, index_base)
        return vals, idx, flags

    def select(self, candidates: torch.Tensor, acq: torch.Tensor, evaluated: torch.Tensor, batch_size: int,
               index_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """select_next_batch on device (reference acquisition.py:116-144).

        Takes the top-(batch+slack) scores, drops rows that equal an evaluated point, keeps the first
        ``batch_size``; the slack grows until enough rows survive or the candidates are exhausted.  If even the
        longest list the top-k kernel supports (BO_MAX_TOPK) consists of evaluated rows only, every evaluated
        candidate is masked out first (``mask_evaluated``) and the list is taken from the masked scores, so the
        result is the reference's for any number of evaluated points.
        Returns (values, global indices) as host arrays, best first; fewer than ``batch_size`` rows only when
        fewer un-evaluated candidates exist (as in the reference).
        """
        n = acq.numel()
        cap = min(n, _lib.BO_MAX_TOPK)
        k = min(n, batch_size + 16)
        masked = False
        while True:
            vals, idx, flags = self.select_listed(candidates, acq, evaluated, k, index_base)
            v = vals.cpu().numpy()
            i = idx.cpu().numpy()
            keep = (flags.cpu().numpy() == 0) & (i >= 0)
            if keep.sum() >= batch_size or (k >= cap and (masked or k >= n)):
                return v[keep][:batch_size], i[keep][:batch_size]
            if k >= cap:
                acq = self.mask_evaluated(acq, candidates, evaluated)
                masked = True
                k = min(n, batch_size + 16)
                continue
            k = min(cap, k * 4)

    def topk_merge(self, vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k of gathered (value, index) pairs with the same total order (after an all-gather)."""
        n_pairs = vals.numel()
        k = int(min(k, n_pairs, _lib.BO_MAX_TOPK))
        ov = torch.empty(k, dtype=_F64, device=self.device)
        oi = torch.empty(k, dtype=torch.int64, device=self.device)
        ws_bytes = self.lib.bo_topk_workspace_bytes(n_pairs, k)
        ws = self.ws.get("topk", ws_bytes, self.device)
        _lib.check(self.lib.bo_topk_merge_f64(_ptr(ov), _ptr(oi), _ptr(vals.contiguous()), _ptr(idx.contiguous()),
                                              n_pairs, k, _ptr(ws), ws_bytes, _stream()))
        return ov, oi


def device_info() -> dict:
    lib = _lib.load()
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    l2, hbm = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(lib.bo_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(l2),
                                  ctypes.byref(hbm)))
    return dict(sm_count=sm.value, cc=(maj.value, mnr.value), l2_bytes=l2.value, hbm_bytes=hbm.value)


class PinnedMirror:
    """Grow-only pinned host tensors keyed by name: D2H targets that do not reallocate per iteration."""

    def __init__(self):
        self._bufs: Dict[str, torch.Tensor] = {}

    def get(self, key: str, shape, dtype=_F64) -> torch.Tensor:
        buf = self._bufs.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self._bufs[key] = buf
        return buf


def hot_path_iteration(gp: DeviceGP, x_vector, y_vector, input_space, prior_mean, prior_variance, length_scales,
                       betas, current_eval: int, batch_size: int, *, mirror: Optional[PinnedMirror] = None,
                       index_base: int = 0, want_host=("mu", "var", "acq")) -> dict:
    """Steps b..h of the reference loop (bayesian_optimization.py:129-207) with HOST buffers in and out.

    Host -> device: x_vector[:n], y_vector[:n], input_space.  Device -> host: the arrays the reference's
    ``state`` dict exposes (mu_objectives, variance_objectives, acquisition_values) plus the selected batch.
    Returns host NumPy views (pinned) under "mu", "var", "acq", the batch rows "x_next", their global
    indices "idx", and the device-side candidate lists used for a multi-GPU merge.
    """
    mirror = mirror or PinnedMirror()
    dev = gp.device
    n = int(current_eval)
    x_dev = to_device(x_vector[:n], _F64, dev)
    y_dev = to_device(y_vector[:n], _F64, dev)
    ready = None
    if isinstance(input_space, torch.Tensor) and input_space.device.type == "cpu" and input_space.is_pinned() \
            and input_space.is_contiguous():
        # the candidate upload (the largest transfer of the step) runs on a copy stream while the factorisation,
        # which does not need the candidates, runs on the caller's stream
        cand_dev = getattr(gp, "_cand_buf", None)
        if cand_dev is None or cand_dev.shape != input_space.shape or cand_dev.dtype != input_space.dtype:
            cand_dev = torch.empty(input_space.shape, dtype=input_space.dtype, device=dev)
            gp._cand_buf = cand_dev
        side = getattr(gp, "_copy_stream", None)
        if side is None:
            side = gp._copy_stream = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))  # earlier readers of the buffer have finished
        with torch.cuda.stream(side):
            cand_dev.copy_(input_space, non_blocking=True)
            ready = side.record_event()
    else:
        cand_dev = to_device(input_space, None, dev)
    gp.fit(x_dev, y_dev, prior_mean, prior_variance, length_scales, n)
    if ready is not None:
        torch.cuda.current_stream(dev).wait_event(ready)
    n_cand = cand_dev.shape[0]
    cache = getattr(gp, "_iter_out", None)
    if cache is None or cache["acq"].numel() != n_cand or cache["mu"].shape[0] != gp.m:
        cache = {k: torch.empty((n_cand,) if k == "acq" else (gp.m, n_cand), dtype=_F64, device=dev)
                 for k in ("mu", "var", "std_mu", "std_var", "ucb", "acq")}
        gp._iter_out = cache
    # One scoring pass over the whole candidate set, then the arrays the reference's `state` exposes go back over
    # PCIe.  (Scoring in slices so that finished slices leave while the next one is computed was measured and is
    # slower: 72.1 vs 67.8 ms per cfg2 step with the FP64 engine, 41 vs 24 ms with the INT8 engine -- the persistent
    # one-CTA-per-SM contraction kernels and the slice-wise copies do not overlap the way the arithmetic suggests.)
    gp.score(cand_dev, betas, out=cache, guard=False)
    host = {key: mirror.get(key, tuple(cache[key].shape)) for key in want_host}
    gp.check_int8_guard(cand_dev)  # INT8 engine only: one sampled cross-check for the whole candidate set
    k = min(n_cand, batch_size + 16)
    vals, idx = gp.topk(cache["acq"], k, index_base)
    flags = gp.match_rows(idx, cand_dev, x_dev, index_base)
    res = {"top_vals_dev": torch.where(flags.bool(), torch.full_like(vals, float("-inf")), vals),
           "top_idx_dev": idx, "device": cache}
    hv = mirror.get("top_vals", (k,))
    hi = mirror.get("top_idx", (k,), torch.int64)
    hf = mirror.get("top_flags", (k,), torch.uint8)
    hv.copy_(vals, non_blocking=True)
    hi.copy_(idx, non_blocking=True)
    hf.copy_(flags, non_blocking=True)
    for key in want_host:
        host[key].copy_(cache[key], non_blocking=True)
        res[key] = host[key].numpy()
    torch.cuda.synchronize(dev)
    keep = (hf.numpy() == 0) & (hi.numpy() >= 0)
    if keep.sum() < batch_size and k < n_cand:  # rare: the slack was eaten by already-evaluated rows
        _, sel = gp.select(cand_dev, cache["acq"], x_dev, batch_size, index_base)
    else:
        sel = hi.numpy()[keep][:batch_size]
    res["idx"] = np.array(sel, dtype=np.int64)
    local = res["idx"] - index_base
    if isinstance(input_space, torch.Tensor):
        res["x_next"] = input_space[torch.from_numpy(local)].cpu().numpy() if len(local) else np.array([])
    else:
        res["x_next"] = np.array([input_space[i] for i in local])
    return res
