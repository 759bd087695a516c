"""Opt-in exact hypervolume improvement, fused with UCB (north star: "UCB and 2-/3-objective HVI become one fused
per-candidate kernel against a sorted Pareto front"): device-prepared fronts (bo_hvi_prepare_f64), the fused
stand-alone pass (bo_acquisition_hvi_f64), the fused scoring epilogue (bo_score_hvi_f64) and the raw-vector entry
bo_hvi_f64 -- all against oracle.gp_oracle.exact_hvi.  PARITY UNPINNED: the reference's "HVI" is sum-UCB
(acquisition.py:104-108), so the oracle here is the builder's own specification, not reference output."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bayesopt_smart_b200 import acquisition as aq
    from bayesopt_smart_b200.engine import DeviceGP, HviFront, to_device

    return dict(aq=aq, DeviceGP=DeviceGP, HviFront=HviFront, to_device=to_device)


def _staircase(rng, n_front, n_extra, m):
    """n_front mutually non-dominated points plus n_extra dominated / duplicate / below-reference / NaN ones."""
    if m == 2:
        t = np.sort(rng.random(n_front))
        front = np.stack([t, 1.0 - t ** 2], axis=1)
    else:
        a, b = rng.random(n_front) * np.pi / 2, rng.random(n_front) * np.pi / 2
        front = np.stack([np.cos(a) * np.cos(b), np.sin(a) * np.cos(b), np.sin(b)], axis=1)
        front = front[orc.pareto_mask_definition(front)]
    extra = front[rng.integers(0, len(front), n_extra)] * rng.uniform(0.2, 0.999, (n_extra, 1))  # dominated
    pts = np.vstack([front, extra, front[:3], np.full((2, m), -5.0)])  # duplicates, points below the reference
    pts[-1, 0] = np.nan
    return pts[rng.permutation(len(pts))], len(front)


@pytest.mark.parametrize("m,n_front,n_extra,n_cand,n_check", [
    (2, 40, 30, 5000, 400),      # ordinary front
    (2, 1500, 600, 5000, 120),   # more than 1024 non-dominated points: no cap for m = 2 (binary search)
    (3, 150, 100, 3000, 60),
    (3, 1400, 100, 600, 5),      # > 1024 live points: swept from global memory
])
def test_fused_ucb_hvi_matches_specification(env, m, n_front, n_extra, n_cand, n_check):
    aq, HviFront, to_device = env["aq"], env["HviFront"], env["to_device"]
    rng = np.random.default_rng(10 * m + n_front)
    pts, live = _staircase(rng, n_front, n_extra, m)
    ref = np.full(m, -0.05)
    front = HviFront(pts, ref)
    got_front = front.points()
    clean = np.maximum(pts[~np.isnan(pts).any(axis=1)], ref)
    want_front = clean[orc.pareto_mask_definition(clean)]
    want_front = want_front[np.lexsort(tuple(-want_front[:, o] for o in reversed(range(m))))]
    assert got_front.shape[0] >= live
    assert np.array_equal(np.unique(got_front, axis=0), np.unique(want_front, axis=0))  # same set of points
    assert np.all(np.diff(got_front[:, 0]) <= 0)                                        # objective 0 descending
    if n_front > 1024:
        assert got_front.shape[0] > 1024
    # candidates straddle the front: some dominated (HVI 0), some beyond it
    prior_mean, prior_var, betas = rng.normal(size=m), rng.uniform(0.5, 2.0, m), rng.uniform(0.5, 2.5, m)
    mu = prior_mean[:, None] + np.sqrt(prior_var)[:, None] * rng.uniform(-0.2, 0.9, (m, n_cand))
    var = prior_var[:, None] * rng.uniform(0.0, 0.08, (m, n_cand))
    ucb, hvi = aq.ucb_and_exact_hvi_device(to_device(mu), to_device(var), prior_mean, prior_var, betas, front)
    ucb_want = (mu - prior_mean[:, None]) / np.sqrt(prior_var)[:, None] + betas[:, None] * np.sqrt(var / prior_var[:, None])
    np.testing.assert_allclose(ucb.cpu().numpy(), ucb_want, rtol=1e-14, atol=1e-15)
    sel = np.linspace(0, n_cand - 1, n_check).astype(int)
    want = orc.exact_hvi(ucb.cpu().numpy()[:, sel].T, pts[~np.isnan(pts).any(axis=1)], ref)
    np.testing.assert_allclose(hvi.cpu().numpy()[sel], want, rtol=1e-10, atol=1e-12)
    h = hvi.cpu().numpy()
    assert (h >= -1e-13).all() and (h == 0).any() and (h > 1e-6).any()
    # the raw-vector entry point (bo_hvi_f64; fronts beyond 1024 points go through the prepared path too)
    raw = aq.exact_hvi_device(ucb, pts[~np.isnan(pts).any(axis=1)], ref)
    np.testing.assert_allclose(raw.cpu().numpy()[sel], want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("engine", ["dmma", "int8"])
@pytest.mark.parametrize("fn,m,d", [("zdt1", 2, 6), ("dtlz2", 3, 8)])
def test_scoring_epilogue_writes_hvi_in_the_same_pass(env, engine, fn, m, d):
    """bo_score_hvi_f64: acq from the fused epilogue == the stand-alone fused pass on the same mu / var, bit for bit
    (same device function, same inputs), and == the specification on a sample."""
    aq, DeviceGP, HviFront = env["aq"], env["DeviceGP"], env["HviFront"]
    n = 300
    x, y, mu0, var0 = orc.make_training_set(fn, n, d, seed=2)
    ls, betas = np.full(m, 0.4), np.full(m, 2.0)
    gp = DeviceGP(variance_engine=engine)
    gp.fit(x, y, mu0, var0, ls, n)
    cand = np.random.default_rng(4).random((20000, d))
    y_std = (y - mu0) / np.sqrt(var0)
    ref = y_std.min(axis=0) - 0.1
    front = HviFront(y_std, ref)
    out = gp.score(cand, betas, want=("mu", "var", "ucb", "acq"), hvi=front)
    plain = gp.score(cand, betas, want=("ucb", "acq"))
    assert torch.equal(out["ucb"], plain["ucb"])
    assert not torch.equal(out["acq"], plain["acq"])  # sum-UCB by default, HVI with a front
    ucb2, hvi2 = aq.ucb_and_exact_hvi_device(out["mu"], out["var"], mu0, var0, betas, front)
    assert torch.equal(ucb2, out["ucb"]) and torch.equal(hvi2, out["acq"])
    sel = np.arange(0, 20000, 20000 // (40 if m == 3 else 300))
    want = orc.exact_hvi(out["ucb"][:, sel].T.cpu().numpy(), y_std, ref)
    np.testing.assert_allclose(out["acq"][sel].cpu().numpy(), want, rtol=1e-10, atol=1e-12)


def test_empty_front_gives_the_box(env):
    aq, HviFront, to_device = env["aq"], env["HviFront"], env["to_device"]
    for m in (2, 3):
        front = HviFront(np.zeros((0, m)), np.zeros(m))
        u = np.abs(np.random.default_rng(m).normal(size=(m, 100))) + 0.1
        _, hvi = aq.ucb_and_exact_hvi_device(to_device(u), to_device(np.zeros_like(u)), np.zeros(m), np.ones(m),
                                             np.zeros(m), front)
        np.testing.assert_allclose(hvi.cpu().numpy(), np.prod(u, axis=0), rtol=1e-14)


def test_bo_loop_with_exact_hvi(env):
    import bayesopt_smart_b200 as bo
    from bayesopt_smart_b200.workloads import toy_function

    np.random.seed(7)
    opt = bo.BayesianOptimization(function=toy_function, bounds=[(0, 300), (0, 300)], n_objectives=2, n_iterations=8,
                                  initial_samples=10, acquisition="exact_hvi")
    opt.optimize()
    assert np.isfinite(opt.acquisition_values).all() and (opt.acquisition_values >= 0).all()
    assert opt.y_vector[:31, 0].max() > 0 and opt.y_vector[:31, 1].max() > -80  # moved towards (150, 150)
