"""GPU parity tests: the CUDA path (through the C ABI / the reference-signature wrappers) against
(a) golden vectors produced by the live reference and (b) the CPU oracle on seeded inputs.

Tolerances: posterior mean / variance in STANDARDISED units,
|d mu|/sqrt(var0) and |d var|/var0 <= tau = max(1e-9, 10*eps*cond(K + 1e-6 I)); elementwise
stages 1e-14 relative; Pareto masks, top-k indices and selected rows bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
CASES = ["gp_float_m2", "gp_float_m3", "gp_intgrid_m3"]


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bayesopt_smart_b200 as p

    return p


def _state(g):
    n = int(g["n"])
    return n, g["x_vector"], g["y_vector"], g["input_space"], g["prior_mean"], g["prior_variance"], \
        g["length_scales"], g["betas"]


def _tau(cond):
    return max(1e-9, 10 * EPS * float(np.max(cond)))


# ------------------------------------------------------------------ function-level drop-ins vs the reference


@pytest.mark.parametrize("case", CASES)
def test_update_k_and_k_star(pkg, golden, case):
    from bayesopt_smart_b200 import numba_kernels as nk

    g = golden(case)
    n, x, y, cand, mu0, var0, ls, betas = _state(g)
    m, total = y.shape[1], x.shape[0]
    k = np.full((m, total, total), -7.0)
    nk.update_k(kernel_matrix=k, x_vector=x, last_eval=0, current_eval=n, prior_variance=var0, length_scales=ls)
    np.testing.assert_allclose(k[:, :n, :n], g["kernel"], rtol=1e-14, atol=0)
    assert np.all(k[:, n:, :] == -7.0) and np.all(k[:, :, n:] == -7.0)  # only [:n,:n] is written
    assert np.array_equal(k[:, :n, :n], np.transpose(k[:, :n, :n], (0, 2, 1)))
    # incremental form: rows/cols [last_eval, current_eval) only
    k2 = np.full((m, total, total), -7.0)
    nk.update_k(k2, x, n - 3, n, var0, ls)
    np.testing.assert_allclose(k2[:, n - 3:n, n - 3:n], g["kernel"][:, n - 3:, n - 3:], rtol=1e-14, atol=0)
    assert np.all(k2[:, : n - 3, :] == -7.0)
    if "k_star" in g:
        ks = np.zeros((m, total, cand.shape[0]))
        nk.update_k_star(k_star=ks, x_vector=x, input_space=cand, last_eval=0, current_eval=n, prior_variance=var0,
                         length_scales=ls)
        np.testing.assert_allclose(ks[:, :n, :], g["k_star"], rtol=1e-14, atol=0)
        assert np.all(ks[:, n:, :] == 0.0)


@pytest.mark.parametrize("case", CASES)
def test_invert_k(pkg, golden, case):
    from bayesopt_smart_b200 import numba_kernels as nk

    g = golden(case)
    n = int(g["n"])
    total = g["x_vector"].shape[0]
    m = g["kernel"].shape[0]
    kbuf = np.zeros((m, total, total))
    kbuf[:, :n, :n] = g["kernel"]
    kinv = nk.invert_k(current_eval=n, kernel_matrix=kbuf)
    assert kinv.shape == (m, n, n)
    scale = np.abs(g["kinv"]).max()
    assert np.abs(kinv - g["kinv"]).max() <= 50 * EPS * g["cond"].max() * scale
    for o in range(m):  # it really is the inverse
        res = kinv[o] @ (g["kernel"][o] + 1e-6 * np.eye(n)) - np.eye(n)
        assert np.abs(res).max() <= 100 * EPS * g["cond"][o]


def test_invert_k_not_positive_definite(pkg):
    from bayesopt_smart_b200 import numba_kernels as nk

    k = np.array([[[1.0, 2.0], [2.0, 1.0]]])
    with pytest.raises(np.linalg.LinAlgError):
        nk.invert_k(2, k)


@pytest.mark.parametrize("case", ["gp_float_m2", "gp_intgrid_m3"])
def test_dense_mean_variance_standardize_ucb(pkg, golden, case):
    """update_mean / update_variance / standardize / update_ucb / update_hypervolume_improvement on the
    reference's own intermediate arrays."""
    from bayesopt_smart_b200 import acquisition as aq
    from bayesopt_smart_b200 import numba_kernels as nk

    g = golden(case)
    n, x, y, cand, mu0, var0, ls, betas = _state(g)
    m, n_cand = g["mu"].shape
    total = x.shape[0]
    ks = np.zeros((m, total, n_cand))
    orc.ref_update_k_star(ks, x, cand, 0, n, var0, ls)
    mu = np.zeros((m, n_cand))
    var = np.zeros((m, n_cand))
    nk.update_mean(mu_objectives=mu, k_star=ks, inverted_kernel_matrix=g["kinv"], y_vector=y, prior_mean=mu0,
                   current_eval=n)
    nk.update_variance(variance_objectives=var, k_star=ks, inverted_kernel_matrix=g["kinv"], prior_variance=var0,
                       current_eval=n)
    tau = _tau(g["cond"])
    for o in range(m):
        assert np.abs(mu[o] - g["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(var[o] - g["var"][o]).max() / var0[o] <= tau
    smu, svar, ucb = np.zeros_like(mu), np.zeros_like(mu), np.zeros_like(mu)
    nk.standardize_objectives(smu, svar, g["mu"], g["var"], mu0, var0)
    np.testing.assert_allclose(smu, g["std_mu"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(svar, g["std_var"], rtol=1e-15, atol=0)
    aq.update_ucb(ucb, g["std_mu"], g["std_var"], betas)
    np.testing.assert_allclose(ucb, g["ucb"], rtol=1e-15, atol=0)
    one = aq.upper_confidence_bound(g["std_mu"][0], g["std_var"][0], float(betas[0]))
    np.testing.assert_allclose(one, g["ucb"][0], rtol=1e-15, atol=0)
    acq = np.zeros(n_cand)
    aq.update_hypervolume_improvement(acq, g["ucb"])
    assert np.array_equal(acq, g["acq"])
    x_next = aq.select_next_batch(cand, g["acq"], x[:n], int(g["batch_size"]))
    assert x_next.dtype == g["x_next"].dtype and np.array_equal(x_next, g["x_next"])


def test_sum_ucb_order(pkg):
    from bayesopt_smart_b200 import acquisition as aq

    acq = np.zeros(2)
    aq.update_hypervolume_improvement(acq, np.array([[1e16, 1.0], [1.0, 2.0], [-1e16, 3.0]]))
    assert acq.tolist() == [0.0, 6.0]


# ------------------------------------------------------------------ fused device path vs the reference


@pytest.mark.parametrize("case", CASES)
def test_fused_path_matches_reference(pkg, golden, case):
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    g = golden(case)
    n, x, y, cand, mu0, var0, ls, betas = _state(g)
    m = y.shape[1]
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n)
    out = gp.score(cand, betas, want=("mu", "var", "std_mu", "std_var", "ucb", "acq"))
    tau = _tau(g["cond"])
    for o in range(m):
        assert np.abs(out["mu"][o].cpu().numpy() - g["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o].cpu().numpy() - g["var"][o]).max() / var0[o] <= tau
        assert np.abs(out["std_mu"][o].cpu().numpy() - g["std_mu"][o]).max() <= tau
        assert np.abs(out["std_var"][o].cpu().numpy() - g["std_var"][o]).max() <= tau
    # UCB takes a square root of the (clamped) variance: the error bound is beta*sqrt(tau) where
    # std_var ~ tau (at training points) and ~tau elsewhere
    bound = sum(tau + betas[o] * np.sqrt(tau) for o in range(m))
    assert np.abs(out["acq"].cpu().numpy() - g["acq"]).max() <= bound
    far = np.all(g["std_var"] > 1e-6, axis=0)
    assert np.abs(out["acq"].cpu().numpy() - g["acq"])[far].max() <= 1e3 * m * tau
    # Bit-exact selection is only claimed where it is provable: every gap of the REFERENCE's ranking that the
    # selection crosses (between consecutive picks, and between the last pick and the runner-up) must exceed twice
    # the largest score difference actually observed between the two implementations -- then no rounding
    # difference can reorder them.  The gap condition is asserted, not assumed.
    acq_dev = out["acq"].cpu().numpy()
    b = int(g["batch_size"])
    seen = {tuple(r) for r in np.asarray(x[:n], dtype=np.float64)}
    order = [i for i in np.argsort(-g["acq"], kind="stable") if tuple(np.asarray(cand[i], dtype=np.float64)) not in seen]
    ranked = g["acq"][order[: b + 1]]
    observed = np.abs(acq_dev - g["acq"]).max()
    assert np.min(-np.diff(ranked)) > 2.0 * observed, (np.diff(ranked), observed)
    vals, idx = gp.select(to_device(cand), out["acq"], to_device(x[:n]), b)
    x_next = np.array([cand[i] for i in idx])
    assert np.array_equal(x_next, g["x_next"])


def test_cfg1_teacher_forced_iterations(pkg, golden):
    """BASELINE config 1 (300x300 int64 grid, toy function): replay the reference's hyper-parameters.
    Iteration 0 has cond ~ 15 -> agreement ~1e-12 and identical x_next."""
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    g = golden("cfg1_trace")
    ranges = [np.arange(0, 300), np.arange(0, 300)]
    cand = np.stack([a.ravel() for a in np.meshgrid(*ranges, indexing="ij")], axis=-1)
    assert cand.dtype == np.int64
    hp = g["hyperparams"][0]
    gp = DeviceGP()
    gp.fit(g["x_vector"], g["y_vector"], g["prior_mean"], hp[2:], hp[:2], 10)
    out = gp.score(cand, g["betas"])
    sub = g["sub_index"]
    for o in range(2):
        assert np.abs(out["mu"][o].cpu().numpy()[sub] - g["mu_sub_0"][o]).max() / np.sqrt(hp[2 + o]) <= 1e-10
        assert np.abs(out["var"][o].cpu().numpy()[sub] - g["var_sub_0"][o]).max() / hp[2 + o] <= 1e-10
    acq = out["acq"].cpu().numpy()
    assert np.abs(acq[sub] - g["acq_sub_0"]).max() <= 1e-6
    _, idx = gp.select(to_device(cand), out["acq"], to_device(g["x_vector"][:10]), 3)
    assert np.array_equal(cand[idx], g["x_next"][0])


def test_powell_fit_matches_reference_trace(pkg, golden):
    """VERDICT r1 #6 / ADVICE r1: teacher-forced Powell parity.  From the reference's recorded initial state of
    BASELINE config 1 (10 LHS points, default length scales, prior variance of the initial y) the GPU-driven
    Powell fit must land on the hyper-parameters the reference's own loop recorded for its first iteration
    (``state["hyperparams"]``, numba_kernels.py:238-321) within Powell's own ``xtol`` (config.py:75, relative
    1e-3).  The length scales agree far better (1e-6); the prior-variance coordinates are flat directions of the
    MLL (it only sees K / prior_variance, numba_kernels.py:195-197), so Powell moves them by rounding noise only
    (reference: 3.5426e7 -> 3.5430e7) and xtol is all that can be asked there.  With the fitted values the
    selected batch must then be the reference's recorded x_next."""
    from bayesopt_smart_b200 import config as cfg
    from bayesopt_smart_b200 import numba_kernels as nk
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    g = golden("cfg1_trace")
    x, y = g["x_vector"].copy(), g["y_vector"].copy()
    ls = np.full(2, cfg.DEFAULT_LENGTH_SCALE, dtype=np.float64)
    pv = g["prior_variance_init"].copy()
    res = nk.optimize_hyperparams_mll(x_vector=x, y_vector=y, kernel_matrix=np.zeros((2, 70, 70)),
                                      prior_mean=g["prior_mean"], prior_variance=pv, length_scales=ls, current_eval=10)
    want = g["hyperparams"][0]
    assert np.array_equal(res.x[:2], ls) and np.array_equal(res.x[2:], pv)  # written in place, like the reference
    np.testing.assert_allclose(ls, want[:2], rtol=1e-6)
    np.testing.assert_allclose(pv, want[2:], rtol=cfg.HYPERPARAM_XTOL)
    # the trajectory continues identically: same batch as the reference picked with ITS fitted values
    ranges = [np.arange(0, 300), np.arange(0, 300)]
    cand = np.stack([a.ravel() for a in np.meshgrid(*ranges, indexing="ij")], axis=-1)
    gp = DeviceGP()
    gp.fit(x, y, g["prior_mean"], pv, ls, 10)
    out = gp.score(cand, g["betas"], want=("acq",))
    _, idx = gp.select(to_device(cand), out["acq"], to_device(x[:10]), 3)
    assert np.array_equal(cand[idx], g["x_next"][0])


def test_cfg1_ill_conditioned_iterations_do_not_fail(pkg, golden):
    """Later cfg1 iterations have cond(K + 1e-6 I) ~ 1e15 (entries ~3.5e7, absolute jitter 1e-6): rounding can
    push a Cholesky pivot below zero although every exact pivot is >= jitter.  The reference's LU inverse
    limps on there (its own posterior is off by 1e-2..1e-1); the CUDA path must not raise: pivots
    within rounding of zero are clamped to the jitter.  Replays all 20 teacher-forced iterations."""
    from bayesopt_smart_b200 import _lib
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    g = golden("cfg1_trace")
    ranges = [np.arange(0, 300), np.arange(0, 300)]
    cand = to_device(np.stack([a.ravel() for a in np.meshgrid(*ranges, indexing="ij")], axis=-1))
    gp = DeviceGP()
    clamped = 0
    for it, n in enumerate(g["iteration"]):
        hp = g["hyperparams"][it]
        gp.fit(g["x_vector"], g["y_vector"], g["prior_mean"], hp[2:], hp[:2], int(n))
        clamped += _lib.load().bo_last_clamped_pivots()
        out = gp.score(cand, g["betas"])
        for o in range(2):
            var = out["var"][o]
            assert torch.isfinite(out["mu"][o]).all() and torch.isfinite(var).all()
            assert var.min().item() >= 1e-10 and var.max().item() <= hp[2 + o] * (1 + 1e-9)
        _, idx = gp.select(cand, out["acq"], to_device(g["x_vector"][: int(n)]), 3)
        assert len(idx) == 3
    # at cond ~ 1e15 some iterations do need the clamp; the count is surfaced (bo_last_clamped_pivots) and small
    print("clamped pivots over the 20 cfg1 iterations:", clamped)
    assert clamped <= 4 * len(g["iteration"]), clamped


# ------------------------------------------------------------------ fused path vs the oracle at larger sizes


@pytest.mark.parametrize("name,n,d,m,ls,n_cand", [
    ("zdt1", 300, 6, 2, 0.3, 5000),     # 3 row blocks (odd count -> unpaired middle block)
    ("zdt1", 1024, 6, 2, 0.3, 4099),    # cfg2 training shape, ragged candidate count
    ("dtlz2", 520, 8, 3, 0.5, 3000),    # 3 objectives, padding 520 -> 640
    ("zdt2", 700, 10, 2, 0.5, 2000),    # d = 10 (cfg3 input width)
    ("zdt1", 130, 6, 2, 0.3, 40000),    # more candidates than one chunk (two chunks + ragged tail)
])
def test_fused_path_matches_oracle(pkg, name, n, d, m, ls, n_cand):
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    x, y, mu0, var0 = orc.make_training_set(name, n, d, seed=0)
    rng = np.random.default_rng(5)
    cand = rng.random((n_cand, d))
    cand[17] = x[3]
    lsv = np.full(m, ls)
    betas = np.full(m, 2.0)
    want = orc.chol_hot_path(x, y, cand, mu0, var0, lsv, betas, n, 3)
    cond = max(np.linalg.cond(want["kernel"][o] + 1e-6 * np.eye(n)) for o in range(m))
    tau = _tau(cond)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, lsv, n)
    out = gp.score(cand, betas, want=("mu", "var", "acq"))
    for o in range(m):
        assert np.abs(out["mu"][o].cpu().numpy() - want["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o].cpu().numpy() - want["var"][o]).max() / var0[o] <= tau
    acq = out["acq"].cpu().numpy()
    far = np.all(want["std_var"] > 1e-6, axis=0)
    assert np.abs(acq - want["acq"])[far].max() <= 1e3 * m * tau
    vals, idx = gp.select(to_device(cand), out["acq"], to_device(x), 3)
    order = orc.ranked_indices(want["acq"])
    gap = want["acq"][order[:5]][:-1] - want["acq"][order[:5]][1:]
    if gap.min() > 1e4 * m * tau:  # only claim bit-exact indices when the oracle's own gaps allow it
        assert np.array_equal(idx, want["idx"])
    # candidate 17 coincides with a training point: variance clamps near the jitter level, never selected
    assert 17 not in idx.tolist()


def test_shard_invariance(pkg):
    """Scoring a candidate set in two shards gives bit-identical numbers (no data-path collective needed)."""
    from bayesopt_smart_b200.engine import DeviceGP

    x, y, mu0, var0 = orc.make_training_set("zdt1", 400, 6, seed=2)
    cand = np.random.default_rng(9).random((3001, 6))
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, [0.3, 0.3], 400)
    full = gp.score(cand, [2.0, 2.0])
    a = gp.score(cand[:1234], [2.0, 2.0])
    b = gp.score(cand[1234:], [2.0, 2.0])
    for k in ("mu", "var", "acq"):
        joined = torch.cat([a[k], b[k]], dim=-1)
        assert torch.equal(joined, full[k]), k


# ------------------------------------------------------------------ selection


def test_topk_matches_total_order(pkg):
    from bayesopt_smart_b200.engine import DeviceGP

    gp = DeviceGP()
    rng = np.random.default_rng(0)
    for n, k in [(5, 3), (8192, 19), (100_003, 64), (1_000_000, 19)]:
        a = rng.normal(size=n)
        a[rng.integers(0, n, size=max(1, n // 50))] = 1.5  # ties
        if n > 100:
            a[7] = np.nan
            a[11] = np.inf
            a[13] = -np.inf
        vals, idx = gp.topk(torch.from_numpy(a).cuda(), k, index_base=1000)
        order = orc.ranked_indices(a)[:k]
        assert np.array_equal(idx.cpu().numpy() - 1000, order), (n, k)
        np.testing.assert_array_equal(vals.cpu().numpy(), a[order])


@pytest.mark.parametrize("kind", ["all_equal", "periodic", "ascending", "descending", "mostly_nan", "top_ties",
                                  "all_nan"])
def test_topk_filtered_scan_adversarial(pkg, kind):
    """Large inputs take the sampled-threshold + streaming-filter path; these inputs stress the threshold
    (massive ties, periodic patterns aligned with grids, NaNs) and must still equal the total order."""
    from bayesopt_smart_b200.engine import DeviceGP

    n = 700_003
    rng = np.random.default_rng(1)
    if kind == "all_equal":
        a = np.full(n, 0.25)
    elif kind == "periodic":
        a = np.tile(np.array([0.0] * 63 + [1.0]), n // 64 + 1)[:n] + 1e-9 * (np.arange(n) % 1000)
    elif kind == "ascending":
        a = np.arange(n, dtype=np.float64)
    elif kind == "descending":
        a = -np.arange(n, dtype=np.float64)
    elif kind == "mostly_nan":
        a = np.full(n, np.nan)
        a[rng.choice(n, 500, replace=False)] = rng.normal(size=500)
    elif kind == "top_ties":
        a = rng.normal(size=n)
        a[rng.choice(n, 5000, replace=False)] = 10.0
    else:
        a = np.full(n, np.nan)
    gp = DeviceGP()
    t = torch.from_numpy(a).cuda()
    for k in (3, 19, 200):
        vals, idx = gp.topk(t, k, index_base=7)
        order = orc.ranked_indices(a)[:k]
        assert np.array_equal(idx.cpu().numpy() - 7, order), (kind, k)
        np.testing.assert_array_equal(vals.cpu().numpy(), a[order])


def test_select_skips_evaluated_rows_and_exhaustion(pkg):
    from bayesopt_smart_b200 import acquisition as aq

    cand = np.arange(40, dtype=np.int64).reshape(20, 2)
    acq = np.linspace(1.0, 0.0, 20)
    ev = cand[[0, 1, 3]].astype(np.float64)
    got = aq.select_next_batch(cand, acq, ev, 3)
    want, _ = orc.ref_select_next_batch(cand, acq, ev, 3)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    # many evaluated rows at the top: slack has to grow
    ev = cand[:19].astype(np.float64)
    got = aq.select_next_batch(cand, acq, ev, 3)
    assert np.array_equal(got, cand[19:20])
    # everything evaluated -> empty result like np.array([])
    got = aq.select_next_batch(cand, acq, cand.astype(np.float64), 3)
    assert got.shape == (0,)


def test_topk_merge_equals_single_shot(pkg):
    from bayesopt_smart_b200.engine import DeviceGP

    gp = DeviceGP()
    a = np.random.default_rng(4).normal(size=50_000)
    a[100] = a[40_000]  # cross-shard tie
    t = torch.from_numpy(a).cuda()
    parts = [gp.topk(t[lo:hi], 8, index_base=lo) for lo, hi in [(0, 12_500), (12_500, 25_000), (25_000, 50_000)]]
    vals = torch.cat([p[0] for p in parts])
    idx = torch.cat([p[1] for p in parts])
    mv, mi = gp.topk_merge(vals, idx, 8)
    sv, si = gp.topk(t, 8)
    assert torch.equal(mi, si) and torch.equal(mv, sv)


# ------------------------------------------------------------------ Pareto


def test_pareto_golden_masks(pkg, golden):
    g = golden("pareto")
    names = sorted(k[:-2] for k in g if k.endswith("_y"))
    for name in names:
        got = pkg.is_pareto_efficient(g[name + "_y"])
        assert got.dtype == bool and np.array_equal(got, g[name + "_mask"]), name
    xs = np.arange(5)[:, None].astype(float)
    px, py = pkg.compute_pareto_front(xs, g["kat_y"])
    assert np.array_equal(px[:, 0], [0, 1, 2, 4])


@pytest.mark.parametrize("n,m", [(3000, 2), (5000, 3), (70_000, 3), (200_000, 2)])
def test_pareto_matches_definition(pkg, n, m):
    rng = np.random.default_rng(n)
    y = rng.normal(size=(n, m))
    y[5] = y[900]
    y[::97] = np.round(y[::97], 1)  # ties
    want = orc.pareto_mask_definition(y) if n <= 5000 else None
    got = pkg.is_pareto_efficient(y)
    if want is None:
        # large sets: check against the definition restricted to the survivors (exact: any dominated
        # point is dominated by an efficient one) and that every dropped point has a dominator there
        front = y[got]
        assert np.array_equal(orc.pareto_mask_definition(front), np.ones(front.shape[0], bool))
        dropped = y[~got][:: max(1, (~got).sum() // 2000)]
        ge = np.all(front[None, :, :] >= dropped[:, None, :], axis=2)
        gt = np.any(front[None, :, :] > dropped[:, None, :], axis=2)
        assert np.all(np.any(ge & gt, axis=1))
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,m", [(65_537, 2), (300_000, 3), (150_000, 4)])
def test_filtered_pareto_pass_equals_the_direct_kernel(pkg, n, m):
    """The in-library large-n pass (sample fronts, thinning, device-side compaction; no host round trips) must give
    bit for bit the mask of the plain n x n kernel -- duplicates, ties and NaN rows included."""
    from bayesopt_smart_b200.engine import to_device
    from bayesopt_smart_b200.pareto import _mask_direct, pareto_mask_device

    rng = np.random.default_rng(n + m)
    y = rng.normal(size=(n, m))
    y[::97] = np.round(y[::97], 1)           # ties
    best = np.argsort(-y.sum(axis=1))[:50]
    y[best[:10] + 1] = y[best[:10]]          # duplicates of strong points: both rows stay efficient
    y[123, 0] = np.nan                       # NaN rows neither dominate nor are dominated
    yd = to_device(y)
    direct = _mask_direct(yd)
    filtered = pareto_mask_device(yd)
    assert torch.equal(direct, filtered)
    assert 0 < int(direct.sum().item()) < n and bool(direct[123].item())


def test_pareto_all_efficient_and_empty(pkg):
    t = np.linspace(0, 1, 4000)
    y = np.stack([t, 1 - t], axis=1)
    assert pkg.is_pareto_efficient(y).all()
    assert pkg.is_pareto_efficient(np.zeros((0, 2))).shape == (0,)


# ------------------------------------------------------------------ MLL


def test_mll_matches_reference(pkg, golden):
    from bayesopt_smart_b200 import numba_kernels as nk

    g = golden("mll")
    n = int(g["n"])
    x, y, mu0 = g["x_vector"], g["y_vector"], g["prior_mean"]
    for s, want in zip(g["settings"], g["mll"]):
        k = np.zeros((2, x.shape[0], x.shape[0]))
        got = nk.compute_mll(x_vector=x, y_vector=y, kernel_matrix=k, prior_mean=mu0, prior_variance=s[2:].copy(),
                             length_scales=s[:2].copy(), current_eval=n)
        assert abs(got - want) <= 1e-8 * max(1.0, abs(want))
        kk = np.zeros_like(k)
        orc.ref_update_k(kk, x, 0, n, s[2:], s[:2])
        np.testing.assert_allclose(k, kk, rtol=1e-14)  # side effect of the reference (:178) kept
    batch = nk.mll_batched(x, y, mu0, g["settings"][:, :2], np.full(4, 1e-8), n)
    np.testing.assert_allclose(batch, g["mll"], rtol=1e-8)


def test_mll_grid_matches_oracle(pkg):
    from bayesopt_smart_b200 import numba_kernels as nk

    x, y, mu0, _ = orc.make_training_set("zdt1", 200, 6, seed=3)
    ls = np.repeat(np.logspace(-1, 0.5, 4), 3)
    jit = np.tile(np.logspace(-8, -2, 3), 4)
    want = orc.mll_grid(x, y, mu0, ls, jit, 200)
    got = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1), jit, 200)
    np.testing.assert_allclose(got, want, rtol=1e-7)


def test_powell_fit_runs_and_improves(pkg):
    from bayesopt_smart_b200 import numba_kernels as nk

    x, y, mu0, var0 = orc.make_training_set("zdt1", 40, 3, seed=4)
    ls = np.array([1.0, 1.0])
    var = var0.copy()
    k = np.zeros((2, 40, 40))
    before = orc.ref_compute_mll(x, y, np.zeros_like(k), mu0, var, ls, 40)
    res = nk.optimize_hyperparams_mll(x, y, k, mu0, var, ls, 40)
    after = orc.ref_compute_mll(x, y, np.zeros_like(k), mu0, var, ls, 40)
    assert after >= before and len(res.x) == 4 and np.allclose(res.x[:2], ls)


# ------------------------------------------------------------------ exact HVI (opt-in; oracle = definition)


@pytest.mark.parametrize("m", [2, 3])
def test_exact_hvi_matches_definition(pkg, m):
    from bayesopt_smart_b200 import acquisition as aq

    rng = np.random.default_rng(m)
    pts = rng.normal(size=(60, m))
    front = pts[orc.pareto_mask_definition(pts)]
    ref = pts.min(axis=0) - 0.5
    u = rng.normal(size=(500, m)) * 1.5
    u[0] = front[0]  # on the front: zero improvement
    u[1] = ref - 1.0  # below the reference point
    want = orc.exact_hvi(u, front, ref)
    got = np.zeros(500)
    aq.update_exact_hypervolume_improvement(got, np.ascontiguousarray(u.T), front, ref)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    assert got[1] == 0.0 and abs(got[0]) <= 1e-12


# ------------------------------------------------------------------ end to end through the public class


def test_bayesian_optimization_class_end_to_end(pkg):
    """Free-running BASELINE config 1 shape (smaller iteration count): callbacks get NumPy state with the
    reference's keys; the optimum of the toy function ([150,150] -> [100, 20]) is found."""
    seen = []

    def toy(xv):
        return np.array([-((xv[0] - 150) ** 2) + 100, -((xv[1] - 150) ** 2) + 20])

    def cb(state):
        assert set(state) == {"iteration", "n_evaluations", "x_vector", "y_vector", "mu_objectives",
                              "variance_objectives", "acquisition_values", "x_next", "hyperparams", "timings"}
        assert set(state["timings"]) == {"hyperparams", "kernels", "acquisition", "eval", "total"}
        assert isinstance(state["mu_objectives"], np.ndarray) and state["mu_objectives"].shape == (2, 90000)
        assert state["x_next"].dtype == np.int64 and len(state["hyperparams"]) == 4
        seen.append(state["iteration"])

    np.random.seed(42)
    opt = pkg.BayesianOptimization(function=toy, bounds=[(0, 300), (0, 300)], n_objectives=2, n_iterations=12,
                                   initial_samples=10, callbacks=[cb, pkg.PerformanceMonitor()])
    opt.optimize()
    assert seen == list(range(10, 46, 3))
    assert opt.n_evaluations == 44  # last_eval + 1 quirk of the reference (:247)
    front = opt.pareto_analysis()
    best = opt.y_vector[:46].max(axis=0)
    assert best[0] >= 100 - 25 and best[1] >= 20 - 25  # within 5 grid steps of both optima
    assert front.shape[1] == 2 and front.shape[0] >= 1
