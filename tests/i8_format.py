"""NumPy restatement of the INT8 engine's number format and layouts (bayesopt_smart_b200/csrc/ozaki.cu).

Test infrastructure shared by tests/test_i8_format_cpu.py (no GPU) and tests/test_gpu_i8.py.  Written from the
format description only: balanced base-256 digits, per-row power-of-two scale for W, fixed 2^46 scale for the
kernel entries, and the canonical K-major shared-memory image both operands are stored in."""
import numpy as np

S = 6          # digit planes
TN = 64        # candidates per tile
KS = 32        # k depth of one MMA
K_SCALE = 2.0 ** 46
BIAS = 0x0000008080808080


def balanced_digits(q):
    """int64 array -> S arrays of digits, most significant first; q == sum_s d[s] * 256**(S-1-s),
    d[s] in [-128, 127] for s > 0."""
    out = [None] * S
    r = q.astype(np.int64).copy()
    for s in range(S - 1, 0, -1):
        d = ((r + 128) & 255) - 128
        out[s] = d
        r = (r - d) >> 8
    out[0] = r
    return out


def balanced_digits_bias_trick(q):
    """The device's way: bytes of (q + BIAS) with the five low bytes XORed by 0x80, read as int8."""
    u = (q.astype(np.int64) + np.int64(BIAS)) ^ np.int64(BIAS)
    out = []
    for s in range(S):
        b = (u >> (8 * (S - 1 - s))) & 255
        out.append(np.where(b >= 128, b - 256, b))
    return out


def w_row_scale(w):
    """Per-row exponent e (max|w| * 128/126 <= 2^e), the scale 2^(e-29) used when recombining and the
    quantisation factor 2^(47-e).  Rows of zeros get scale 0."""
    mx = np.abs(w).max(axis=1)
    nz = mx > 0
    e = np.zeros(w.shape[0], dtype=np.int64)
    e[nz] = np.frexp(mx[nz] * (128.0 / 126.0))[1]
    return np.where(nz, np.ldexp(1.0, e - 29), 0.0), np.where(nz, np.ldexp(1.0, 47 - e), 0.0)


def quantize_w(w):
    ws, qs = w_row_scale(w)
    return balanced_digits(np.rint(w * qs[:, None]).astype(np.int64)), ws


def quantize_kstar(kt):
    """kt = exp(-0.5 |x - c|^2 / ls^2) in [0, 1]."""
    return balanced_digits(np.rint(kt * K_SCALE).astype(np.int64))


def contraction(wd, ws, kd, pairs_max=S - 1):
    """sum_i V_i^2 per candidate, V = (row scale) * (b0 + 2^-24 b1) from the exact digit-pair sums
    acc_g = sum_{s+t=g} Wd[s] @ Kd[t]^T, b0 = (acc0*256 + acc1)*256 + acc2, b1 likewise from acc3..5.
    wd[s]: (rows, k) int64, kd[t]: (cands, k) int64.  Multiply by prior_var**2 for the device's output."""
    acc = [np.zeros((wd[0].shape[0], kd[0].shape[0]), dtype=np.int64) for _ in range(2 * S - 1)]
    for s in range(S):
        for t in range(S):
            if s + t <= pairs_max:
                acc[s + t] += wd[s] @ kd[t].T
    assert max(int(np.abs(a).max()) for a in acc) < 2 ** 31
    f = ws[:, None]
    if pairs_max == S - 1:  # the device's recombination, operation for operation
        b0 = (acc[0] * 256 + acc[1]) * 256 + acc[2]
        b1 = (acc[3] * 256 + acc[4]) * 256 + acc[5]
        v = b0.astype(np.float64) * f + b1.astype(np.float64) * (f / 16777216.0)
    else:  # any other pair set: V = f 2^16 sum_g 256^-g acc_g in extended precision
        v = sum(a.astype(np.longdouble) * np.longdouble(256.0) ** (2 - g) for g, a in enumerate(acc)) * f
    return np.asarray((v * v).sum(axis=0), dtype=np.float64)


def planes_from_image(buf, rows, nk):
    """Device image [k-step][plane][row group of 8][k half][row in group][16 bytes] -> (S, rows, nk*32) digits."""
    a = np.frombuffer(buf, dtype=np.int8) if not isinstance(buf, np.ndarray) else buf.view(np.int8)
    a = a.reshape(nk, S, rows // 8, 2, 8, 16).transpose(1, 2, 4, 0, 3, 5).reshape(S, rows, nk * KS)
    return a.astype(np.int64)


def unpack_wpack(wp, npad):
    """DMMA-packed W (bayesopt_smart_b200/csrc/common.cuh) -> dense lower-triangular (npad, npad)."""
    r = np.arange(npad)[:, None]
    k = np.arange(npad)[None, :]
    ib, rr, kt, kk = r >> 7, r & 127, k >> 4, k & 15
    wm, i, g = rr >> 6, (rr >> 3) & 7, rr & 7
    sp, q, t = kk >> 3, (kk >> 2) & 1, kk & 3
    off = (8 * ib * (ib + 1) // 2 + kt) * 2048 + (((wm * 8 + i) * 2 + sp) * 32 + (4 * g + t)) * 2 + q
    mask = np.broadcast_to(k < (ib + 1) * 128, (npad, npad))
    w = np.zeros((npad, npad))
    w[mask] = wp[np.broadcast_to(off, (npad, npad))[mask]]
    return np.tril(w)
