"""CPU checks of the INT8 engine's arithmetic design (no GPU): the digit format is an exact representation,
the device's bias/XOR trick produces the same digits, and the truncated set of 21 digit pairs reproduces the
FP64 quadratic form far inside the 1e-9 parity tolerance on the reference's kind of data."""
import numpy as np
import scipy.linalg as sla

from oracle import gp_oracle as orc
from tests import i8_format as f8


def test_balanced_digits_roundtrip_and_range():
    rng = np.random.default_rng(0)
    q = rng.integers(-(2 ** 47) + 2 ** 40, 2 ** 47 - 2 ** 40, size=20000, dtype=np.int64)
    q[:4] = [0, 1, -1, 2 ** 46]
    d = f8.balanced_digits(q)
    recon = sum(d[s] * 256 ** (f8.S - 1 - s) for s in range(f8.S))
    assert np.array_equal(recon, q)
    for s in range(1, f8.S):
        assert d[s].min() >= -128 and d[s].max() <= 127
    assert np.abs(d[0]).max() <= 127  # fits a signed byte given the 126/128 head-room of the row scale


def test_bias_xor_trick_equals_balanced_digits():
    rng = np.random.default_rng(1)
    q = rng.integers(-(2 ** 47) + 2 ** 40, 2 ** 47 - 2 ** 40, size=20000, dtype=np.int64)
    a, b = f8.balanced_digits(q), f8.balanced_digits_bias_trick(q)
    for s in range(f8.S):
        assert np.array_equal(a[s], b[s]), s


def test_row_scale_leaves_headroom():
    rng = np.random.default_rng(2)
    w = np.tril(rng.standard_normal((64, 64)) * np.exp(rng.uniform(-20, 20, size=(64, 1))))
    w[5] = 0.0
    ws, qs = f8.w_row_scale(w)
    q = np.rint(w * qs[:, None])
    assert np.abs(q).max() <= 2 ** 47 * 126 / 128 + 1
    assert ws[5] == 0.0 and qs[5] == 0.0
    np.testing.assert_array_equal(ws[ws > 0] * qs[qs > 0], 2.0 ** 18)  # 2^(e-29) * 2^(47-e)


def _problem(n, d, ls, n_cand, seed=0):
    x, y, mu0, var0 = orc.make_training_set("zdt1", n, d, seed=seed)
    cand = np.random.default_rng(seed + 1).random((n_cand, d))
    sq = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    k = var0[0] * np.exp(-0.5 * sq / ls ** 2) + 1e-6 * np.eye(n)
    w = sla.solve_triangular(np.linalg.cholesky(k), np.eye(n), lower=True)
    sqs = ((x[:, None, :] - cand[None, :, :]) ** 2).sum(-1)
    kt = np.exp(-0.5 * sqs / ls ** 2)
    return w, kt, var0[0], np.linalg.cond(k)


def test_truncated_digit_pairs_reproduce_the_quadratic_form():
    """max |q_i8 - q_exact| / var0 for 21 pairs (s + t <= 5): far below 1e-9, also at cond ~ 1e7."""
    for n, d, ls, bound in [(256, 6, 0.3, 2e-11), (256, 6, 0.6, 2e-11), (384, 3, 0.3, 2e-11)]:
        w, kt, var0, cond = _problem(n, d, ls, 300)
        v = w.astype(np.longdouble) @ (var0 * kt).astype(np.longdouble)
        q_exact = np.asarray((v * v).sum(axis=0), dtype=np.float64)
        wd, ws = f8.quantize_w(w)
        kd = f8.quantize_kstar(kt.T)
        q_i8 = f8.contraction(wd, ws, kd) * var0 ** 2
        err = np.abs(q_i8 - q_exact).max() / var0
        assert err < bound, (n, d, ls, cond, err)
        # every digit pair (36) is exact to FP64 rounding level: the error above is the truncation alone
        q_full = f8.contraction(wd, ws, kd, pairs_max=2 * f8.S) * var0 ** 2
        assert np.abs(q_full - q_exact).max() / var0 < 5e-13 * max(1.0, cond / 1e6)


def test_image_layout_roundtrip():
    rng = np.random.default_rng(3)
    rows, nk = 64, 3
    dig = rng.integers(-128, 128, size=(f8.S, rows, nk * f8.KS), dtype=np.int64)
    img = np.zeros(nk * f8.S * rows * f8.KS, dtype=np.int8)
    for ks in range(nk):
        for s in range(f8.S):
            for r in range(rows):
                for k in range(f8.KS):
                    off = ((ks * f8.S + s) * (rows // 8) + r // 8) * 256 + (k // 16) * 128 + (r % 8) * 16 + k % 16
                    img[off] = dig[s, r, ks * f8.KS + k]
    assert np.array_equal(f8.planes_from_image(img, rows, nk), dig)


def _row_block(t, r, nsplit):
    """Python restatement of oz_row_block (ozaki.cu): serpentine dealing of row blocks to the CTAs of a tile."""
    return t * nsplit + ((nsplit - 1 - r) if (t & 1) else r)


def test_serpentine_row_block_dealing_is_a_balanced_partition():
    for nb in (1, 2, 5, 8, 16, 32, 33, 128):
        for nsplit in (1, 2, 3, 4, 7, 8):
            if nsplit > nb:
                continue
            seen, ksteps = [], []
            for r in range(nsplit):
                mine, t = [], 0
                while _row_block(t, r, nsplit) < nb:
                    mine.append(_row_block(t, r, nsplit))
                    t += 1
                assert mine, (nb, nsplit, r)          # every CTA of a tile has work (nsplit <= nb)
                seen += mine
                ksteps.append(sum(4 * (ib + 1) for ib in mine))  # k-steps of 32: row block ib spans 4 (ib + 1)
            assert sorted(seen) == list(range(nb))    # every row block exactly once
            # loads differ by at most the k-steps of the two largest blocks (an odd tail round)
            assert max(ksteps) - min(ksteps) <= 8 * nb, (nb, nsplit, ksteps)
            assert all(k % 4 == 0 for k in ksteps)    # the two MMA issuers rely on even k-step counts per block


def test_accumulator_bound_at_the_size_limit():
    """|acc_g| <= (pairs in g) * 128 * 128 * N must stay below 2^31 for N = 16384 (OZ_MAX_N) and the widest group."""
    assert 6 * 128 * 128 * 16384 < 2 ** 31
    assert 6 * 128 * 128 * (16384 + 128 * 44) >= 2 ** 31  # ... and not by a wide margin: the limit is real
