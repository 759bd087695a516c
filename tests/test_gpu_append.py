"""SURVEY 8(f)2 -- incremental factor update (bo_gp_append_f64 behind DeviceGP.fit): the factor extended by the
new rows must equal the factor rebuilt from scratch up to rounding, <= 10 eps cond relative to the largest entry
(the reference always rebuilds: update_k / invert_k with last_eval = 0, bayesian_optimization.py:129-142), and the
predictions that follow must stay inside the parity tolerance against the CPU oracle."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    return dict(lib=_lib.load(), _lib=_lib, DeviceGP=DeviceGP, to_device=to_device)


def _rel(a, b):
    return float((a - b).abs().max().item() / b.abs().max().item())


@pytest.mark.parametrize("fn,n_old,steps,d,m,ls,cond", [
    ("zdt1", 1000, [3, 3, 3, 15], 6, 2, 0.3, 1.3e4),   # cfg2's training shape, batch 3 (and one larger batch)
    ("dtlz2", 5, [3, 32], 8, 3, 0.5, 1e3),              # first iterations of a run, 3 objectives, max batch
    ("zdt1", 4090, [3, 3], 6, 2, 0.3, 3.5e6),           # north-star size, last block before the padding ends
])
def test_appended_factor_equals_full_refit(env, fn, n_old, steps, d, m, ls, cond):
    DeviceGP = env["DeviceGP"]
    n_total = n_old + sum(steps)
    x, y, mu0, var0 = orc.make_training_set(fn, n_total, d, seed=0)
    lsv, betas = np.full(m, ls), np.full(m, 2.0)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, lsv, n_old)
    assert gp.last_fit == "full"
    n = n_old
    for b in steps:
        n += b
        gp.fit(x, y, mu0, var0, lsv, n)
        assert gp.last_fit == "append", (n, gp.last_fit)
        ref = DeviceGP()
        ref.fit(x, y, mu0, var0, lsv, n, incremental=False)
        assert ref.last_fit == "full"
        tol = 10 * EPS * cond
        assert _rel(gp.wpack, ref.wpack) <= tol
        assert _rel(gp.alpha, ref.alpha) <= max(tol, 1e-12)
    # predictions from the extended factor: same tolerance as any other fit, against the CPU oracle
    cand = np.random.default_rng(5).random((3000, d))
    cand[7] = x[n - 1]  # the newest training point: variance must collapse there
    out = gp.score(cand, betas, want=("mu", "var", "acq"))
    want = orc.chol_hot_path(x, y, cand, mu0, var0, lsv, betas, n, 3)
    tau = max(1e-9, 10 * EPS * cond)
    for o in range(m):
        assert np.abs(out["mu"][o].cpu().numpy() - want["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o].cpu().numpy() - want["var"][o]).max() / var0[o] <= tau
        assert out["var"][o][7].item() <= 1e-4 * var0[o]


def test_append_conditions_and_fallbacks(env):
    DeviceGP, to_device = env["DeviceGP"], env["to_device"]
    n0, d, m = 120, 4, 2
    x, y, mu0, var0 = orc.make_training_set("zdt1", 200, d, seed=3)
    ls = np.full(m, 0.4)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n0)
    gp.fit(x, y, mu0, var0, ls, n0 + 3)
    assert gp.last_fit == "append"
    gp.fit(x, y, mu0, var0, ls * (1 + 1e-15), n0 + 6)      # hyper-parameters not bit-identical -> rebuild
    assert gp.last_fit == "full"
    gp.fit(x, y, mu0, var0, ls * (1 + 1e-15), n0 + 7)
    assert gp.last_fit == "append"
    gp.fit(x, y, mu0, var0, ls * (1 + 1e-15), n0 + 10)     # 127 -> 130 leaves the 128-row padding -> rebuild
    assert gp.last_fit == "full" and gp.n == n0 + 10
    x2 = x.copy()
    x2[3, 0] += 1e-9                                        # an OLD row changed -> rebuild
    gp.fit(x2, y, mu0, var0, ls * (1 + 1e-15), n0 + 12)
    assert gp.last_fit == "full"
    gp.fit(x2, y, mu0, var0, ls * (1 + 1e-15), n0 + 12 + 40)  # more than BO_MAX_APPEND rows at once -> rebuild
    assert gp.last_fit == "full"
    gp.fit(x2, y, mu0, var0, ls * (1 + 1e-15), n0 + 12 + 43, incremental=False)
    assert gp.last_fit == "full"
    # device-resident inputs take the same decisions (prefix compared on the device)
    xd, yd = to_device(x), to_device(y)
    g2 = DeviceGP(variance_engine="int8")
    g2.fit(xd, yd, mu0, var0, ls, n0)
    g2.fit(xd, yd, mu0, var0, ls, n0 + 3)
    assert g2.last_fit == "append"
    ref = DeviceGP(variance_engine="int8")
    ref.fit(x, y, mu0, var0, ls, n0 + 3)
    cand = np.random.default_rng(1).random((500, d))
    a = g2.score(cand, np.full(m, 2.0), want=("var",))["var"]
    b = ref.score(cand, np.full(m, 2.0), want=("var",))["var"]
    assert float((a - b).abs().max().item()) <= 1e-10 * float(var0.max())  # the digit planes were re-quantised


def test_append_abi_argument_checks(env):
    lib, _lib = env["lib"], env["_lib"]
    z = torch.zeros(1 << 16, dtype=torch.float64, device="cuda")
    hp = (ctypes.c_double * 2)(1.0, 1.0)
    args = lambda n_old, n_new: (z.data_ptr(), z.data_ptr(), z.data_ptr(), 4, z.data_ptr(), 2, n_old, n_new, 4, 2,  # noqa: E731
                                 hp, hp, hp, 1e-6, z.data_ptr(), 0, None)
    assert lib.bo_gp_append_f64(*args(100, 140)) == _lib.BO_ERR_INVALID    # more than BO_MAX_APPEND rows
    assert lib.bo_gp_append_f64(*args(127, 130)) == _lib.BO_ERR_INVALID    # leaves the padding
    assert lib.bo_gp_append_f64(*args(100, 100)) == _lib.BO_ERR_INVALID    # nothing to add
    assert lib.bo_gp_append_f64(*args(100, 103)) == _lib.BO_ERR_WORKSPACE  # workspace_bytes = 0


def test_bo_loop_reuses_the_factor_within_tolerance(env):
    """optimize(): default tolerance 0 -> every iteration rebuilds (hyper-parameters move), the trace is the plain
    one; a generous tolerance -> the factor is extended instead, and the run still finds the toy optimum."""
    import bayesopt_smart_b200 as bo
    from bayesopt_smart_b200 import bayesian_optimization as bom
    from bayesopt_smart_b200.workloads import toy_function

    def run(tol):
        np.random.seed(42)
        opt = bo.BayesianOptimization(function=toy_function, bounds=[(0, 300), (0, 300)], n_objectives=2,
                                      n_iterations=12, initial_samples=10, hyperparam_tolerance=tol)
        opt.optimize()
        return opt, list(bom.LAST_RUN_INFO["fit_modes"])

    plain, modes0 = run(0.0)
    assert set(modes0) == {"full"}
    reuse, modes1 = run(0.5)
    assert modes1[0] == "full" and "append" in modes1
    assert plain.y_vector[:46].max(axis=0).tolist() == [100.0, 20.0]
    assert reuse.y_vector[:46, 0].max() >= 75.0 and reuse.y_vector[:46, 1].max() >= -5.0  # close to (100, 20)
