"""GPU parity tests of the INT8 tensor-core variance engine (variance_engine="int8", ozaki.cu) through the
C ABI: digit planes bit-exact against the NumPy restatement of the format, the tcgen05 contraction exact
against the integer contraction of the digits it was given, and the whole scoring pass against the CPU oracle
with the same tolerance as the FP64 DMMA engine: |d mu|/sqrt(var0), |d var|/var0 <= max(1e-9, 10*eps*cond)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc
from tests import i8_format as f8

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    return dict(lib=_lib.load(), _lib=_lib, DeviceGP=DeviceGP, to_device=to_device)


def _fit(env, n, d, m, ls, engine="int8", fn=None, seed=0):
    fn = fn or {1: "zdt1", 2: "zdt1", 3: "dtlz2", 4: "dtlz2"}[m]
    x, y, mu0, var0 = orc.make_training_set(fn, n, d, seed=seed) if m in (2, 3) else _mk(n, d, m, seed)
    gp = env["DeviceGP"](variance_engine=engine)
    gp.fit(x, y, mu0, var0, np.full(m, ls), n)
    return gp, x, y, mu0, var0


def _mk(n, d, m, seed):
    """training sets for objective counts the oracle's generators do not produce directly (1 and 4)"""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.stack([np.sin(3.0 * x.sum(1) + o) + 0.1 * o * x[:, 0] for o in range(m)], axis=1)
    return x, y, y.mean(0), np.maximum(y.var(0), 1e-3)


def test_digit_planes_and_contraction_are_exact(env):
    lib, _lib = env["lib"], env["_lib"]
    n, d, m, n_cand, ls = 200, 6, 2, 150, 0.3
    gp, x, y, mu0, var0 = _fit(env, n, d, m, ls)
    torch.cuda.synchronize()
    npad, nb, nk = lib.bo_npad(n), lib.bo_npad(n) // 128, lib.bo_npad(n) // 32
    wp = gp.wpack.cpu().numpy().reshape(m, -1)
    wq = gp.wq.cpu().numpy().reshape(m, -1)
    wsc = gp.wscale.cpu().numpy()[: m * npad].reshape(m, npad)
    wdig = []
    for o in range(m):
        w = f8.unpack_wpack(wp[o], npad)
        w[n:, :] = 0.0
        w[:, n:] = 0.0
        want, ws = f8.quantize_w(w)
        got = np.zeros((f8.S, npad, npad), dtype=np.int64)
        for ib in range(nb):
            nks, blk0 = 4 * (ib + 1), 2 * ib * (ib + 1)
            got[:, ib * 128:(ib + 1) * 128, : nks * 32] = f8.planes_from_image(
                wq[o][blk0 * 24576:(blk0 + nks) * 24576], 128, nks)
        for s in range(f8.S):
            assert np.array_equal(got[s], want[s]), (o, s)
        assert np.array_equal(wsc[o], ws)
        wdig.append(got)

    rng = np.random.default_rng(1)
    cand = rng.random((n_cand, d))
    tiles = ((n_cand + f8.TN - 1) // f8.TN + 3) // 4 * 4
    kq = torch.zeros(m * tiles * npad * f8.S * f8.TN, dtype=torch.uint8, device="cuda")
    meandot = torch.zeros(m * tiles * f8.TN, dtype=torch.float64, device="cuda")
    cand_dev = env["to_device"](cand)
    _, pv = _lib.host_doubles(var0, m)
    _, pl = _lib.host_doubles(np.full(m, ls), m)
    _lib.check(lib.bo_i8_kstar_digits(kq.data_ptr(), meandot.data_ptr(), cand_dev.data_ptr(), 0, d, n_cand,
                                      gp.x.data_ptr(), gp.x.stride(0), n, d, m, gp.alpha.data_ptr(), pv, pl, None))
    torch.cuda.synchronize()
    kqh = kq.cpu().numpy().reshape(m, tiles, -1)
    kdig = []
    for o in range(m):
        kt = np.exp(-0.5 * ((x[:n, None, :] - cand[None, :, :]) ** 2).sum(-1) / ls ** 2)
        got = np.concatenate([f8.planes_from_image(kqh[o, t], f8.TN, nk) for t in range(tiles)], axis=1)
        recon = sum(got[s].astype(np.float64) * 256.0 ** (f8.S - 1 - s) for s in range(f8.S)) / f8.K_SCALE
        # half a unit of the last digit plus the 1-ulp difference between the device exp and NumPy's
        assert np.abs(recon[:n_cand, :n].T - kt).max() <= 2.0 ** -47 + 4 * EPS
        assert got[0].min() >= 0 and got[0].max() <= 65
        kdig.append(got)

    for nsplit in (1, nb):
        q_dev = torch.zeros(m * nsplit * tiles * f8.TN, dtype=torch.float64, device="cuda")
        _lib.check(lib.bo_i8_sumsq(q_dev.data_ptr(), gp.wq.data_ptr(), gp.wscale.data_ptr(), kq.data_ptr(), n, m,
                                   n_cand, nsplit, pv, None))
        torch.cuda.synchronize()
        got = q_dev.cpu().numpy().reshape(m, nsplit, tiles * f8.TN).sum(1)
        for o in range(m):
            want = f8.contraction(wdig[o], wsc[o], kdig[o]) * var0[o] ** 2
            # integer sums are exact; the FP64 recombination differs from NumPy's only by FMA contraction
            np.testing.assert_allclose(got[o][:n_cand], want[:n_cand], rtol=2e-14, atol=0)


CASES = [  # n, d, m, ls, n_cand
    (1, 2, 1, 0.5, 70), (10, 2, 2, 0.4, 333), (127, 3, 2, 0.3, 64), (128, 6, 2, 0.3, 65), (129, 6, 3, 0.5, 1000),
    (300, 10, 2, 0.5, 4097), (513, 8, 4, 0.5, 500), (640, 16, 2, 1.0, 129), (1024, 6, 2, 0.3, 3000),
]


@pytest.mark.parametrize("n,d,m,ls,n_cand", CASES)
def test_score_matches_oracle(env, n, d, m, ls, n_cand):
    gp, x, y, mu0, var0 = _fit(env, n, d, m, ls)
    cand = np.random.default_rng(n + d).random((n_cand, d))
    betas = np.full(m, 2.0)
    out = gp.score(cand, betas, want=("mu", "var", "std_mu", "std_var", "ucb", "acq"))
    torch.cuda.synchronize()
    lsv = np.full(m, ls)
    fit = orc.chol_fit(x, y, mu0, var0, lsv, n)
    mu_w, var_w = orc.chol_predict(fit, x, cand, mu0, var0, lsv, n)
    cond = max(np.linalg.cond(var0[o] * np.exp(-0.5 * ((x[:n, None] - x[None, :n]) ** 2).sum(-1) / ls ** 2)
                              + 1e-6 * np.eye(n)) for o in range(m))
    tau = max(1e-9, 10 * EPS * cond)
    mu_g, var_g = out["mu"].cpu().numpy(), out["var"].cpu().numpy()
    for o in range(m):
        assert np.abs(mu_g[o] - mu_w[o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(var_g[o] - var_w[o]).max() / var0[o] <= tau
    # the elementwise stages behind the contraction are the same kernels as in the DMMA path
    smu = (mu_g - mu0[:, None]) / np.sqrt(var0)[:, None]
    svar = var_g / var0[:, None]
    ucb = smu + betas[:, None] * np.sqrt(np.abs(svar))
    np.testing.assert_allclose(out["std_mu"].cpu().numpy(), smu, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(out["ucb"].cpu().numpy(), ucb, rtol=1e-14, atol=1e-15)
    acq = np.zeros(n_cand)
    for o in range(m):
        acq = acq + out["ucb"][o].cpu().numpy()
    assert np.array_equal(out["acq"].cpu().numpy(), acq)  # sequential sum from 0.0, bit for bit


def test_ill_conditioned_stays_inside_tau(env):
    """cond(K + 1e-6 I) ~ 1e7: the tolerance widens with eps*cond, the digit truncation does not."""
    n, d, m, ls = 512, 6, 2, 0.5
    gp, x, y, mu0, var0 = _fit(env, n, d, m, ls)
    gp2, *_ = _fit(env, n, d, m, ls, engine="dmma")
    cand = np.random.default_rng(5).random((2000, d))
    a = gp.score(cand, np.full(m, 2.0), want=("var",))["var"].cpu().numpy()
    b = gp2.score(cand, np.full(m, 2.0), want=("var",))["var"].cpu().numpy()
    for o in range(m):
        assert np.abs(a[o] - b[o]).max() / var0[o] <= 5e-11


def test_int64_grid_and_multiple_chunks(env):
    """int64 Cartesian grid (the reference's input_space dtype) long enough for several chunks; the result of a
    candidate does not depend on where it sits in the chunk / tile / cluster."""
    n, d, m, ls = 64, 2, 2, 40.0
    x, y, mu0, var0 = _mk(n, d, m, 3)
    x = np.round(x * 300.0)
    gp = env["DeviceGP"](variance_engine="int8")
    gp.fit(x, y, mu0, var0, np.full(m, ls), n)
    ax = np.arange(450, dtype=np.int64)
    grid = np.stack([g.ravel() for g in np.meshgrid(ax, ax, indexing="ij")], axis=-1)  # 202 500 rows
    betas = np.full(m, 1.0)
    full = gp.score(grid, betas, want=("mu", "var", "acq"))
    lo, hi = 77_777, 77_777 + 5_001
    part = gp.score(grid[lo:hi].astype(np.float64), betas, want=("mu", "var", "acq"))
    torch.cuda.synchronize()
    for key in ("mu", "var", "acq"):
        assert torch.equal(full[key][..., lo:hi], part[key]), key
    fit = orc.chol_fit(x, y, mu0, var0, np.full(m, ls), n)
    sel = np.random.default_rng(0).choice(grid.shape[0], 2000, replace=False)
    mu_w, var_w = orc.chol_predict(fit, x, grid[sel].astype(np.float64), mu0, var0, np.full(m, ls), n)
    for o in range(m):
        assert np.abs(full["var"][o].cpu().numpy()[sel] - var_w[o]).max() / var0[o] <= 1e-9


def test_engines_pick_the_same_batch_on_cfg2(env):
    """BASELINE cfg2 (ZDT1 d=6, N=1024, 10^6-point grid): both engines agree to 1e-10 of the prior variance and
    select the same batch."""
    n, d, m, ls = 1024, 6, 2, 0.3
    axes = [np.linspace(0.0, 1.0, 10)] * d
    grid = np.ascontiguousarray(np.stack([g.ravel() for g in np.meshgrid(*axes, indexing="ij")], axis=-1))
    cand = env["to_device"](grid)
    betas = np.full(m, 2.0)
    res = {}
    for eng in ("dmma", "int8"):
        gp, x, y, mu0, var0 = _fit(env, n, d, m, ls, engine=eng)
        out = gp.score(cand, betas, want=("var", "acq"))
        _, idx = gp.select(cand, out["acq"], env["to_device"](x), 3)
        res[eng] = (out["var"].clone(), out["acq"].clone(), idx)
    dv = (res["int8"][0] - res["dmma"][0]).abs().amax(dim=1).cpu().numpy() / var0
    assert dv.max() <= 1e-10, dv
    assert np.array_equal(res["int8"][2], res["dmma"][2])


def test_size_limit_is_an_error_not_a_fallback(env):
    lib, _lib = env["lib"], env["_lib"]
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    rc = lib.bo_i8_quantize_w(buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 16385, 1, None)
    assert rc == _lib.BO_ERR_INVALID
    assert b"16384" in lib.bo_last_error()
    with pytest.raises(ValueError):
        env["DeviceGP"](variance_engine="fp32")


def test_peak_probe_reports_a_plausible_rate(env):
    tops = ctypes.c_double()
    env["_lib"].check(env["lib"].bo_i8_peak_tops(ctypes.byref(tops), 0.01, None))
    assert 1500.0 < tops.value < 5200.0, tops.value  # nominal dense int8 on B200: 4.5 POP/s


def test_optimize_runs_with_the_int8_engine(env):
    """The reference-signature loop (BayesianOptimization.optimize) on the toy 2-D problem with the INT8 engine."""
    from bayesopt_smart_b200 import BayesianOptimization
    from bayesopt_smart_b200.pareto import is_pareto_efficient

    def toy(xv):
        return np.array([-(xv[0] - 20.0) ** 2 - (xv[1] - 10.0) ** 2, -(xv[0] - 5.0) ** 2 - (xv[1] - 25.0) ** 2])

    bo = BayesianOptimization(toy, [(0, 30), (0, 30)], n_objectives=2, n_iterations=3, initial_samples=8,
                              batch_size=3, variance_engine="int8")
    bo.optimize()
    assert np.isfinite(bo.y_vector[: 8 + 9]).all()
    assert is_pareto_efficient(bo.y_vector[: 8 + 9]).any()
    assert bo.variance_objectives.shape[1] == 900 and (bo.variance_objectives >= 1e-10).all()


def test_parity_edge_and_full_size_suites_pass_with_the_int8_engine_as_default(env):
    """Every other GPU test file (golden / oracle parity, edge shapes, BASELINE-size properties, the 2-rank loop)
    re-run with BO_VARIANCE_ENGINE=int8, i.e. with the INT8 engine behind every DeviceGP() of those tests."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join("tests", f) for f in ("test_gpu_parity.py", "test_gpu_edges.py", "test_gpu_full_size.py",
                                                "test_gpu_multi.py")]
    env2 = dict(os.environ, BO_VARIANCE_ENGINE="int8")
    r = subprocess.run([sys.executable, "-m", "pytest", *files, "-m", "gpu", "-q", "-x"], cwd=root, env=env2,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_randomised_shapes_are_deterministic_and_within_tau(env):
    """tools/oz_stress.py: random (n, d, m, candidates, length scale) cases; every case is scored twice with the INT8
    engine (bit-identical) and compared with the FP64 engine (tolerance max(1e-9, 10 eps cond))."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join("tools", "oz_stress.py"), "40", "7"], cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_sampled_guard_passes_and_trips(env):
    """the INT8 engine cross-checks one candidate in `stride` against the FP64 engine on
    every score() and fails loudly -- no fallback -- when the difference exceeds the tolerance."""
    DeviceGP, _lib = env["DeviceGP"], env["_lib"]
    n, d, m = 700, 6, 2
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=0)
    ls, betas = np.full(m, 0.3), np.full(m, 2.0)
    cand = np.random.default_rng(2).random((300_000, d))
    gp = DeviceGP(variance_engine="int8")               # default: tolerance 1e-9, one candidate in 4096
    gp.fit(x, y, mu0, var0, ls, n)
    out = gp.score(cand, betas, want=("var", "acq"))
    assert gp.last_guard_worst is not None and 0.0 < gp.last_guard_worst < 1e-10
    # the sampled numbers are those of the main pass: re-derive the worst difference from the full arrays
    ref = DeviceGP()
    ref.fit(x, y, mu0, var0, ls, n)
    full = ref.score(cand, betas, want=("var",))
    dv = ((out["var"] - full["var"]).abs() / torch.tensor(var0, device="cuda")[:, None]).max().item()
    assert gp.last_guard_worst <= dv <= 1e-10
    # the tolerance the sample was held to: max(1e-9, 10 eps cond_upper), cond_upper = n (var0 + jitter) |W|_F^2 --
    # an upper bound of cond(K + jitter I) = 1.3e4 here, so the tolerance stays within two decades of the 1e-9 floor
    assert 1e-9 <= gp.last_guard_tolerance <= 1e-7
    # a damaged INT8 factor (row scales off by 1e-4): caught at the default tolerance, FP64 engine untouched
    gp.wscale[: 2 * 128] *= 1.0 + 1e-4
    with pytest.raises(_lib.Int8GuardError):
        gp.score(cand, betas, want=("acq",))
    assert gp.last_guard_worst > gp.last_guard_tolerance >= 1e-9
    # switched off explicitly: no check, no error
    off = DeviceGP(variance_engine="int8", int8_guard_tol=0.0)
    off.fit(x, y, mu0, var0, ls, n)
    off.wscale[: 2 * 128] *= 1.0 + 1e-4
    off.score(cand, betas, want=("acq",))
    assert off.last_guard_worst is None
