"""BASELINE.json configs at FULL size on one B200, checked through size-independent properties and oracle
spot checks (the oracle cannot run these sizes in full): shard/merge invariance of the top-k, a strided sample
against the CPU oracle, exclusion of evaluated rows, Pareto front closure, MLL determinism.
Multi-GPU configs are exercised as one rank's shard (the path has no data-path collective)."""
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


def _spot_check(gp, out, cand_dev, x, y, mu0, var0, ls, n, m, count, tau):
    sel = torch.arange(0, cand_dev.shape[0], max(1, cand_dev.shape[0] // count), device="cuda")[:count]
    cs = cand_dev[sel].cpu().numpy()
    fit = orc.chol_fit(x, y, mu0, var0, ls, n)
    mu_o, var_o = orc.chol_predict(fit, x, cs, mu0, var0, ls, n)
    for o in range(m):
        assert np.abs(out["mu"][o][sel].cpu().numpy() - mu_o[o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o][sel].cpu().numpy() - var_o[o]).max() / var0[o] <= tau


def test_cfg2_full_grid(pkg):
    """cfg2: ZDT1 d=6, N=1024, the full 10^6-point grid linspace(0,1,10)^6, 2 objectives, 1 GPU."""
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    n, d, m = 1024, 6, 2
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=0)
    axes = [np.linspace(0.0, 1.0, 10)] * d
    cand = np.stack([g.ravel() for g in np.meshgrid(*axes, indexing="ij")], axis=-1)
    cand[123456] = x[7]  # one grid point replaced by an evaluated point
    ls, betas = np.full(m, 0.3), np.full(m, 2.0)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n)
    cd = to_device(cand)
    out = gp.score(cd, betas, want=("mu", "var", "acq"))
    _spot_check(gp, out, cd, x, y, mu0, var0, ls, n, m, 1200, 1e-9)
    assert torch.isfinite(out["acq"]).all()
    assert out["var"].min().item() >= 1e-10
    for o in range(m):
        assert out["var"][o].max().item() <= var0[o] * (1 + 1e-12)
        # at the evaluated point the variance collapses to ~jitter level and the mean interpolates y
        assert out["var"][o][123456].item() <= 1e-4 * var0[o]
        assert abs(out["mu"][o][123456].item() - y[7, o]) <= 1e-4 * np.sqrt(var0[o])
    # top-k: single shot == merge of 4 shards' lists (what the multi-GPU path does)
    k = 19
    sv, si = gp.topk(out["acq"], k)
    parts = [gp.topk(out["acq"][lo:lo + 250_000], k, index_base=lo) for lo in range(0, 1_000_000, 250_000)]
    mv, mi = gp.topk_merge(torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), k)
    assert torch.equal(si, mi) and torch.equal(sv, mv)
    a = out["acq"].cpu().numpy()
    assert np.array_equal(si.cpu().numpy(), orc.ranked_indices(a)[:k])
    assert np.all(np.diff(sv.cpu().numpy()) <= 0)
    # sharded scoring is bit-identical to the single pass
    half = gp.score(cd[500_000:], betas, want=("acq",))
    assert torch.equal(half["acq"], out["acq"][500_000:])
    _, idx = gp.select(cd, out["acq"], to_device(x), 3)
    assert 123456 not in idx.tolist() and len(set(idx.tolist())) == 3


def test_cfg3_shard_n4096_d10(pkg):
    """cfg3: ZDT2 d=10, N=4096, random candidates; one rank's work at reduced candidate count (the per-candidate
    arithmetic does not depend on M), oracle spot check of 200 candidates."""
    from bayesopt_smart_b200.engine import DeviceGP

    n, d, m = 4096, 10, 2
    x, y, mu0, var0 = orc.make_training_set("zdt2", n, d, seed=0)
    ls, betas = np.full(m, 0.5), np.full(m, 2.0)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n)
    g = torch.Generator(device="cuda").manual_seed(0)
    cd = torch.rand(100_000, d, dtype=torch.float64, device="cuda", generator=g)
    out = gp.score(cd, betas, want=("mu", "var", "acq"))
    _spot_check(gp, out, cd, x, y, mu0, var0, ls, n, m, 200, 1e-9)  # cond ~ 1e5 at these hyper-parameters
    assert torch.isfinite(out["acq"]).all()


def test_cfg4_shard_three_objectives_hvi_pareto(pkg):
    """cfg4: DTLZ2 d=8, N=2048, 3 objectives: scores, exact 3-objective HVI and Pareto filter of the UCB vectors."""
    from bayesopt_smart_b200.acquisition import exact_hvi_device
    from bayesopt_smart_b200.engine import DeviceGP
    from bayesopt_smart_b200.pareto import pareto_mask_device

    n, d, m = 2048, 8, 3
    x, y, mu0, var0 = orc.make_training_set("dtlz2", n, d, seed=0)
    ls, betas = np.full(m, 0.5), np.full(m, 2.0)
    gp = DeviceGP()
    gp.fit(x, y, mu0, var0, ls, n)
    g = torch.Generator(device="cuda").manual_seed(1)
    cd = torch.rand(200_000, d, dtype=torch.float64, device="cuda", generator=g)
    out = gp.score(cd, betas, want=("mu", "var", "ucb", "acq"))
    cond = max(np.linalg.cond(orc.chol_fit(x, y, mu0, var0, ls, n)["kernel"][o] + 1e-6 * np.eye(n)) for o in range(m))
    _spot_check(gp, out, cd, x, y, mu0, var0, ls, n, m, 300, max(1e-9, 10 * EPS * cond))
    y_std = (y - mu0) / np.sqrt(var0)
    front = y_std[orc.pareto_mask_definition(y_std)]
    ref = y_std.min(axis=0) - 0.1
    hv = exact_hvi_device(out["ucb"], front, ref)
    sel = np.arange(0, 200_000, 1000)
    want = orc.exact_hvi(out["ucb"][:, sel].T.cpu().numpy(), front, ref)
    np.testing.assert_allclose(hv[sel].cpu().numpy(), want, rtol=1e-10, atol=1e-12)
    rows = out["ucb"].T.contiguous()
    mask = pareto_mask_device(rows).bool()
    fr = rows[mask].cpu().numpy()
    assert np.array_equal(orc.pareto_mask_definition(fr), np.ones(fr.shape[0], bool))  # front is non-dominated
    dropped = rows[~mask][::97].cpu().numpy()
    ge = np.all(fr[None, :, :] >= dropped[:, None, :], axis=2)
    gt = np.any(fr[None, :, :] > dropped[:, None, :], axis=2)
    assert np.all(np.any(ge & gt, axis=1))  # every dropped point has a dominator on the front


def test_pareto_8m_points(pkg):
    from bayesopt_smart_b200.pareto import pareto_mask_device

    g = torch.Generator(device="cuda").manual_seed(3)
    yv = torch.randn(8_000_000, 3, dtype=torch.float64, device="cuda", generator=g)
    mask = pareto_mask_device(yv).bool()
    front = yv[mask].cpu().numpy()
    assert 10 < front.shape[0] < 5000
    assert orc.pareto_mask_definition(front).all()
    drop = yv[~mask][::4001].cpu().numpy()
    ge = np.all(front[None, :, :] >= drop[:, None, :], axis=2)
    gt = np.any(front[None, :, :] > drop[:, None, :], axis=2)
    assert np.all(np.any(ge & gt, axis=1))
    # idempotence: the front of the front is the front
    assert pareto_mask_device(yv[mask].contiguous()).all()


def test_cfg5_full_sweep(pkg):
    """cfg5: 256 (length scale, jitter) settings at N=4096, d=6: finite, deterministic, and the jitter=1e-8
    settings equal compute_mll (oracle) -- one of them is checked (6 s of CPU)."""
    from bayesopt_smart_b200 import numba_kernels as nk

    n, d, m = 4096, 6, 2
    x, y, mu0, _ = orc.make_training_set("zdt1", n, d, seed=0)
    ls = np.repeat(np.logspace(-1, 0.5, 16), 16)
    jit = np.tile(np.logspace(-8, -2, 16), 16)
    vals = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1), jit, n)
    assert vals.shape == (256,) and np.isfinite(vals).all()
    again = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1)[:32], jit[:32], n)
    assert np.array_equal(again, vals[:32])  # bit-reproducible, independent of the batch composition
    s = 5 * 16  # length scale 0.316, jitter 1e-8
    want = orc.ref_compute_mll(x, y, np.zeros((m, n, n)), mu0, np.ones(m), np.full(m, ls[s]), n)
    assert abs(vals[s] - want) <= 1e-7 * abs(want)


@pytest.mark.parametrize("engine", ["dmma", "int8"])
@pytest.mark.parametrize("case", ["gp_n1024_d6", "gp_n4096_d6", "gp_n4096_d10_zdt2", "gp_n2048_d8_dtlz2"])
def test_reference_fixtures_at_baseline_scale(pkg, golden, case, engine):
    """both variance engines against outputs of the LIVE reference at the cfg2 training size
    (N = 1024, M = 4096), the north-star size (N = 4096, M = 1024), cfg3's shape (ZDT2, d = 10, N = 4096) and cfg4's
    (DTLZ2, d = 8, N = 2048, three objectives) -- not against the builder's own oracle.
    Tolerance max(1e-9, 10 eps cond) in standardised units; the selected batch must be the reference's, and that
    claim is only made after asserting that the reference's ranking gaps exceed twice the observed difference."""
    from bayesopt_smart_b200.engine import DeviceGP, to_device
    from tests.test_oracle_golden import baseline_scale_inputs

    g = golden(case)
    n, x, y, cand, mu0, var0, ls, betas = baseline_scale_inputs(g)
    m = y.shape[1]
    gp = DeviceGP(variance_engine=engine)
    gp.fit(x, y, mu0, var0, ls, n)
    cd = to_device(cand)
    out = gp.score(cd, betas, want=("mu", "var", "std_mu", "std_var", "ucb", "acq"))
    tau = max(1e-9, 10 * EPS * float(g["cond"].max()))
    for o in range(m):
        assert np.abs(out["mu"][o].cpu().numpy() - g["mu"][o]).max() / np.sqrt(var0[o]) <= tau
        assert np.abs(out["var"][o].cpu().numpy() - g["var"][o]).max() / var0[o] <= tau
        assert np.abs(out["std_mu"][o].cpu().numpy() - g["std_mu"][o]).max() <= tau
        assert np.abs(out["std_var"][o].cpu().numpy() - g["std_var"][o]).max() <= tau
    acq = out["acq"].cpu().numpy()
    far = np.all(g["std_var"] > 1e-6, axis=0)  # away from training points the square root does not amplify
    assert np.abs(acq - g["acq"])[far].max() <= 1e3 * m * tau
    b = int(g["batch_size"])
    seen = {tuple(r) for r in x[:n]}
    order = [i for i in np.argsort(-g["acq"], kind="stable") if tuple(cand[i]) not in seen]
    ranked = g["acq"][order[: b + 1]]
    assert np.min(-np.diff(ranked)) > 2.0 * np.abs(acq - g["acq"]).max()
    _, idx = gp.select(cd, out["acq"], to_device(x), b)
    assert np.array_equal(cand[idx], g["x_next"])


def test_mll_and_pareto_against_reference_outputs_at_baseline_sizes(pkg, golden):
    """GPU batched MLL against compute_mll of the LIVE reference on cfg2's and cfg5's training sets (N = 1024 / 4096,
    jitter = CHOLESKY_JITTER), and the dominance kernel against the reference's masks at n = 2048."""
    from bayesopt_smart_b200 import numba_kernels as nk
    from bayesopt_smart_b200.workloads import make_training_set

    g = golden("mll_large")
    for n in (1024, 4096):
        x, y, mu0, _ = make_training_set("zdt1", n, 6, seed=0)
        ls = g[f"n{n}_length_scales"]
        vals = nk.mll_batched(x, y, mu0, np.stack([ls, ls], axis=1), np.full(len(ls), 1e-8), n)
        # length scale 1.0 with jitter 1e-8 is a correlation matrix of condition ~1e11: the fit term y^T R^-1 y then
        # carries eps * cond ~ 1e-5 of itself, i.e. the MLL agrees to ~1e-8..1e-7 relative between any two Cholesky
        # implementations (measured 1.9e-8); the well-conditioned settings agree to 1e-8 and better
        for got, want, scale in zip(vals, g[f"n{n}_mll"], ls):
            assert abs(got - want) <= (1e-8 if scale < 0.9 else 1e-6) * abs(want), (n, scale, got, want)
    p = golden("pareto_large")
    assert np.array_equal(pkg.is_pareto_efficient(p["cloud_y"]), p["cloud_mask"])
    _, yd, _, _ = make_training_set("dtlz2", 2048, 8, seed=0)
    assert np.array_equal(pkg.is_pareto_efficient(yd), p["dtlz2_mask"])


def test_selection_when_more_than_1024_top_rows_were_evaluated(pkg):
    """the listed top-k is capped at BO_MAX_TOPK = 1024; when every listed row is an evaluated point the
    selection masks the evaluated candidates exhaustively (bo_mask_evaluated_f64) and still returns the reference's
    batch (acquisition.py:134-144: first rows of the ranking that are not evaluated)."""
    from bayesopt_smart_b200.engine import DeviceGP, to_device

    rng = np.random.default_rng(12)
    n_cand, d = 300_000, 4
    cand = rng.random((n_cand, d))
    acq = rng.normal(size=n_cand)
    order = np.argsort(-acq)
    evaluated = cand[order[:1500]]           # the 1500 best candidates were all evaluated already
    gp = DeviceGP()
    vals, idx = gp.select(to_device(cand), to_device(acq), to_device(evaluated), 5)
    assert idx.tolist() == order[1500:1505].tolist()
    assert np.array_equal(vals, acq[order[1500:1505]])
    _, want_idx = orc.ref_select_next_batch(cand, acq, evaluated, 5)
    assert idx.tolist() == list(want_idx)


def test_exhaustive_exclusion_with_int64_grid_candidates(pkg):
    """The same fallback on the reference's own candidate type: an int64 Cartesian grid compared with float64
    evaluated rows by value (acquisition.py:139), more than 1024 evaluated points at the top of the ranking."""
    from bayesopt_smart_b200.engine import DeviceGP, grid_candidates, to_device

    bounds = [(0, 80), (0, 60), (0, 50)]
    cand_dev = grid_candidates(bounds)
    cand = cand_dev.cpu().numpy()
    assert cand.dtype == np.int64 and cand.shape == (240_000, 3)
    acq = np.random.default_rng(4).normal(size=cand.shape[0])
    order = np.argsort(-acq)
    evaluated = cand[order[:1100]].astype(np.float64)
    vals, idx = DeviceGP().select(cand_dev, to_device(acq), to_device(evaluated), 4)
    assert idx.tolist() == order[1100:1104].tolist()
