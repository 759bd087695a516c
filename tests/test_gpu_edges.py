"""Edge cases of the CUDA path against the oracle: extreme shapes (one training point, one candidate,
d = 1..16, m = 1..4, sizes just around the 128-row / 128-candidate tile boundaries), int64 candidates,
non-contiguous host inputs, candidates fewer than the batch, NaN scores in the ranking."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bayesopt_smart_b200 as p

    return p


def _problem(n, d, m, n_cand, seed, ls=0.6):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    w = rng.normal(size=(d, m))
    y = np.sin(2.0 * x @ w) + 0.05 * rng.normal(size=(n, m))
    mu0 = y.mean(axis=0)
    var0 = np.maximum(y.var(axis=0), 0.05)
    cand = rng.random((n_cand, d))
    return x, y, mu0, var0, cand, np.full(m, ls * np.sqrt(d) / 2), np.linspace(1.0, 2.5, m)


def _check(pkg, x, y, mu0, var0, cand, ls, betas, n, batch=3):
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    m = y.shape[1]
    want = orc.chol_hot_path(x, y, cand, mu0, var0, ls, betas, n, batch)
    cond = max(np.linalg.cond(want["kernel"][o] + 1e-6 * np.eye(n)) for o in range(m))
    tau = max(1e-9, 10 * EPS * cond)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n)
    out = gp.score(cand, betas, want=("mu", "var", "std_mu", "std_var", "ucb", "acq"))
    for o in range(m):
        assert np.abs(out["mu"][o].cpu().numpy() - want["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o].cpu().numpy() - want["var"][o]).max() / var0[o] <= tau
    far = np.all(want["std_var"] > 1e-6, axis=0)
    if far.any():
        assert np.abs(out["acq"].cpu().numpy() - want["acq"])[far].max() <= 1e3 * m * tau
    _, idx = gp.select(to_device(cand), out["acq"], to_device(x[:n]), batch)
    assert len(idx) == min(batch, len(want["idx"]))
    return out, want, idx


@pytest.mark.parametrize("n,d,m,n_cand", [
    (1, 1, 1, 1),          # the smallest possible problem
    (1, 3, 2, 5),
    (2, 1, 4, 130),        # 4 objectives, 1-D inputs
    (127, 5, 2, 127), (128, 5, 2, 128), (129, 5, 2, 129),   # around the row / candidate tile size
    (255, 7, 3, 257), (257, 9, 1, 255),
    (64, 11, 2, 300), (64, 12, 2, 300), (40, 13, 2, 300), (40, 16, 2, 300),  # every K* dimension bucket
    (384, 2, 2, 1000), (500, 4, 3, 77),
])
def test_shapes(pkg, n, d, m, n_cand):
    x, y, mu0, var0, cand, ls, betas = _problem(n, d, m, n_cand, seed=n * 31 + d)
    _check(pkg, x, y, mu0, var0, cand, ls, betas, n)


def test_buffers_larger_than_current_eval_and_strided_inputs(pkg):
    """The reference passes preallocated (T, d) buffers with only the first n rows valid; also accept
    non-contiguous views (column slices) and Fortran-ordered candidates."""
    x, y, mu0, var0, cand, ls, betas = _problem(90, 4, 2, 500, seed=5)
    xbig = np.full((120, 6), 99.0)
    xbig[:90, 1:5] = x
    ybig = np.full((120, 3), -5.0)
    ybig[:90, :2] = y
    candf = np.asfortranarray(cand)
    out, want, idx = _check(pkg, xbig[:, 1:5], ybig[:, :2], mu0, var0, candf, ls, betas, 90)
    assert np.array_equal(idx, want["idx"])


def test_int64_candidates_and_exclusion(pkg):
    from bayesopt_smart_b200 import acquisition as aq

    rng = np.random.default_rng(8)
    ranges = [np.arange(0, 9), np.arange(0, 7), np.arange(0, 5)]
    cand = np.stack([g.ravel() for g in np.meshgrid(*ranges, indexing="ij")], axis=-1)
    assert cand.dtype == np.int64
    pick = rng.choice(cand.shape[0], 20, replace=False)
    x = cand[pick].astype(np.float64)
    y = np.stack([-(x ** 2).sum(1), x[:, 0] - x[:, 2]], axis=1)
    mu0, var0 = y.mean(0), y.var(0)
    ls, betas = np.array([2.0, 3.0]), np.array([1.0, 2.0])
    out, want, idx = _check(pkg, x, y, mu0, var0, cand, ls, betas, 20, batch=4)
    assert np.array_equal(idx, want["idx"])
    assert not set(idx.tolist()) & set(pick.tolist())
    got = aq.select_next_batch(cand, out["acq"].cpu().numpy(), x, 4)
    assert got.dtype == np.int64 and np.array_equal(got, want["x_next"])


def test_fewer_candidates_than_batch(pkg):
    from bayesopt_smart_b200 import acquisition as aq

    cand = np.array([[0.1, 0.2], [0.5, 0.5]])
    got = aq.select_next_batch(cand, np.array([0.3, 0.9]), np.zeros((0, 2)), 5)
    assert np.array_equal(got, cand[[1, 0]])


def test_nan_and_inf_scores_rank_last_and_first(pkg):
    from bayesopt_smart_b200.engine import DeviceGP

    gp = DeviceGP()
    a = np.array([0.5, np.nan, np.inf, -np.inf, 0.5, -0.0, 0.0, 2.0])
    vals, idx = gp.topk(torch.from_numpy(a).cuda(), 8)
    assert idx.cpu().tolist() == [2, 7, 0, 4, 5, 6, 3, 1]  # inf, 2, ties by index, -0.0 == 0.0 by index, -inf, NaN
    assert np.isnan(vals.cpu().numpy()[-1])


def test_pareto_edge_inputs(pkg):
    assert pkg.is_pareto_efficient(np.array([[1.0, 2.0]])).tolist() == [True]
    same = np.ones((300, 3))
    assert pkg.is_pareto_efficient(same).all()  # duplicates are all kept (pareto.py semantics)
    y = np.array([[np.nan, np.nan], [1.0, 1.0], [0.0, 0.0], [np.inf, -np.inf], [-np.inf, np.inf]])
    assert np.array_equal(pkg.is_pareto_efficient(y), orc.ref_is_pareto_efficient_loop(y))
    one = np.arange(10.0)[:, None]  # a single objective: only the maximum survives
    assert pkg.is_pareto_efficient(one).tolist() == [False] * 9 + [True]
    four = np.random.default_rng(0).normal(size=(400, 4))
    assert np.array_equal(pkg.is_pareto_efficient(four), orc.pareto_mask_definition(four))
    with pytest.raises(ValueError):
        pkg.is_pareto_efficient(np.zeros((3, 5)))


def test_refit_with_growing_training_set_reuses_engine(pkg):
    """BO loop pattern: the same DeviceGP is refitted with n growing by the batch size each iteration."""
    from bayesopt_smart_b200.engine import DeviceGP

    x, y, mu0, var0, cand, ls, betas = _problem(140, 3, 2, 400, seed=11)
    gp = DeviceGP()
    for n in (10, 13, 127, 128, 131, 140):
        gp.fit(x, y, mu0, var0, ls, n)
        out = gp.score(cand, betas)
        fit = orc.chol_fit(x, y, mu0, var0, ls, n)
        mu_o, var_o = orc.chol_predict(fit, x, cand, mu0, var0, ls, n)
        cond = max(np.linalg.cond(fit["kernel"][o] + 1e-6 * np.eye(n)) for o in range(2))
        tau = max(1e-9, 10 * EPS * cond)
        for o in range(2):
            assert np.abs(out["mu"][o].cpu().numpy() - mu_o[o]).max() / np.sqrt(var0[o]) <= tau
            assert np.abs(out["var"][o].cpu().numpy() - var_o[o]).max() / var0[o] <= tau


def test_error_paths(pkg):
    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200.engine import DeviceGP

    gp = DeviceGP()
    with pytest.raises(_lib.BoError):
        gp.score(np.zeros((4, 2)), [1.0])  # score before fit
    x, y, mu0, var0, cand, ls, betas = _problem(20, 3, 2, 50, seed=1)
    gp.fit(x, y, mu0, var0, ls, 20)
    with pytest.raises(ValueError):
        gp.score(np.zeros((4, 5)), betas)  # wrong candidate width
    with pytest.raises(TypeError):
        gp.score(np.zeros((4, 3), dtype=np.float32), betas)
    with pytest.raises(_lib.BoError):
        gp.fit(np.zeros((5, 17)), np.zeros((5, 1)), [0.0], [1.0], [1.0], 5)  # d > BO_MAX_DIMS
    with pytest.raises(_lib.BoError):
        gp.fit(np.zeros((5, 2)), np.zeros((5, 5)), np.zeros(5), np.ones(5), np.ones(5), 5)  # m > BO_MAX_OBJECTIVES
    # duplicated training points: K + 1e-6 I is still positive definite -> no error, finite output
    xd = np.repeat(x[:5], 4, axis=0)
    yd = np.repeat(y[:5], 4, axis=0)
    gp.fit(xd, yd, mu0, var0, ls, 20)
    out = gp.score(cand, betas)
    assert torch.isfinite(out["mu"]).all() and torch.isfinite(out["var"]).all()


def test_device_grid_generator_matches_meshgrid(pkg):
    """bo_grid_i64 / engine.grid_candidates: the reference's candidate set (bayesian_optimization.py:338-340),
    whole and in shards, bit for bit."""
    from bayesopt_smart_b200.engine import grid_candidates

    for bounds in ([(0, 300), (0, 300)], [(-3, 4), (10, 13), (0, 5)], [(0, 10)] * 6, [(5, 6), (0, 7)], [(2, 9)]):
        axes = np.meshgrid(*[np.arange(lo, hi) for lo, hi in bounds], indexing="ij")
        want = np.stack([a.ravel() for a in axes], axis=-1)
        got = grid_candidates(bounds)
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
        total = want.shape[0]
        for lo, hi in ((0, 1), (total // 3, total // 3 + min(1000, total - total // 3)), (total - 1, total), (5, 5)):
            if lo <= hi <= total:
                assert np.array_equal(grid_candidates(bounds, lo, hi - lo).cpu().numpy(), want[lo:hi])
    from bayesopt_smart_b200._lib import BoError

    with pytest.raises(BoError):
        grid_candidates([(0, 4), (3, 3)])


def test_optimize_with_replaced_input_space_uses_it(pkg):
    """A caller that swaps `input_space` for its own candidate list gets that list scored (not the regenerated grid)."""
    def toy(xv):
        return np.array([-(xv[0] - 7.0) ** 2, -(xv[1] - 3.0) ** 2])

    bo = pkg.BayesianOptimization(toy, [(0, 12), (0, 12)], n_objectives=2, n_iterations=2, initial_samples=6,
                                  batch_size=2)
    keep = bo.input_space[(bo.input_space[:, 0] % 2 == 0)]
    bo.input_space = keep
    for name in ("mu_objectives", "variance_objectives", "std_mu_objectives", "std_variance_objectives", "ucb"):
        setattr(bo, name, np.zeros((2, len(keep))))
    bo.acquisition_values = np.zeros(len(keep))
    bo.optimize()
    new_rows = bo.x_vector[6:10]
    assert np.all(new_rows[:, 0] % 2 == 0)


@pytest.mark.parametrize("engine", ["dmma", "int8"])
def test_host_buffer_iteration_slices_equal_single_shot(pkg, engine):
    """engine.hot_path_iteration (the end-to-end call bench.py times): pinned host buffers in, host arrays out -- they
    must be bit for bit those of DeviceGP.score on device-resident inputs, and the batch the one DeviceGP.select
    picks; scoring the same set in column slices of wider output arrays (strided outputs) changes nothing either."""
    from bayesopt_smart_b200.engine import DeviceGP, PinnedMirror, hot_path_iteration, to_device

    n, d, m, n_cand = 200, 6, 2, 700_001  # more than two slices of 4 * 4 * SMs * 128 candidates, ragged tail
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=0)
    ls, betas = np.full(m, 0.3), np.full(m, 2.0)
    cand = np.random.default_rng(9).random((n_cand, d))
    cand[77] = x[3]
    gp = DeviceGP(variance_engine=engine)
    res = hot_path_iteration(gp, torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(),
                             torch.from_numpy(cand).pin_memory(), mu0, var0, ls, betas, n, 3, mirror=PinnedMirror())
    ref = DeviceGP(variance_engine=engine)
    ref.fit(x, y, mu0, var0, ls, n)
    cd = to_device(cand)
    out = ref.score(cd, betas, want=("mu", "var", "acq"))
    for key in ("mu", "var", "acq"):
        assert np.array_equal(res[key], out[key].cpu().numpy()), key
    _, idx = ref.select(cd, out["acq"], to_device(x), 3)
    assert np.array_equal(res["idx"], idx) and 77 not in res["idx"].tolist()
    assert np.array_equal(res["x_next"], cand[idx])
    if engine == "int8":
        assert gp.last_guard_worst is not None and gp.last_guard_worst < 1e-10  # one check for the whole set
    # slice-wise scoring into column slices of the full arrays: a candidate's numbers do not depend on the slicing
    full = {k: torch.empty_like(v) for k, v in out.items()}
    for s0 in range(0, n_cand, 250_000):
        s1 = min(n_cand, s0 + 250_000)
        ref.score(cd[s0:s1], betas, out={k: (v[s0:s1] if v.dim() == 1 else v[:, s0:s1]) for k, v in full.items()},
                  guard=False)
    for key in ("mu", "var", "acq"):
        assert torch.equal(full[key], out[key]), key
