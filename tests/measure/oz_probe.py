"""GPU probe of the INT8 (Ozaki) variance engine, stage by stage, against NumPy restatements of the digit
format and of the exact integer contraction.  Usage: python tests/measure/oz_probe.py [n] [n_cand] [d] [m]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesopt_smart_b200 import _lib  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from oracle import gp_oracle as orc  # noqa: E402

from tests.i8_format import S, balanced_digits as balanced, planes_from_image, unpack_wpack  # noqa: E402


def planes_from_buffer(buf, rows, nk):
    return planes_from_image(buf, rows, nk)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    m = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    lib = _lib.load()
    x, y, mu0, var0 = orc.make_training_set("zdt1" if m == 2 else "dtlz2", n, d, seed=0)
    rng = np.random.default_rng(1)
    cand = rng.random((n_cand, d))
    ls = np.full(m, 0.3)
    betas = np.full(m, 2.0)
    gp = DeviceGP(variance_engine="int8")
    gp.fit(x, y, mu0, var0, ls, n)
    torch.cuda.synchronize()
    npad = lib.bo_npad(n)
    nb = npad // 128
    nk = npad // 32
    wp = gp.wpack.cpu().numpy().reshape(m, -1)
    wq = gp.wq.cpu().numpy().reshape(m, -1)
    wsc = gp.wscale.cpu().numpy()[: m * npad].reshape(m, npad)
    ok = True
    Wd = []
    for o in range(m):
        W = unpack_wpack(wp[o], npad)
        W[n:, :] = 0.0
        W[:, n:] = 0.0
        mx = np.abs(W).max(1)
        e = np.zeros(npad, dtype=np.int64)
        nz = mx > 0
        e[nz] = np.frexp(mx[nz] * (128.0 / 126.0))[1]
        ws_ref = np.where(nz, np.ldexp(1.0, e - 29), 0.0)
        q = np.rint(W * np.where(nz, np.ldexp(1.0, 47 - e), 0.0)[:, None]).astype(np.int64)
        dig = balanced(q)
        # GPU planes: per row block ib, k-steps 0..4(ib+1)-1
        got = np.zeros((S, npad, npad), dtype=np.int64)
        for ib in range(nb):
            nks = 4 * (ib + 1)
            blk0 = 2 * ib * (ib + 1)
            buf = wq[o][blk0 * 24576:(blk0 + nks) * 24576]
            got[:, ib * 128:(ib + 1) * 128, : nks * 32] = planes_from_buffer(buf, 128, nks)
        bad = sum(int((got[s] != dig[s]).sum()) for s in range(S))
        print(f"[W digits] obj {o}: mismatching digits {bad}, scale mismatch {int((wsc[o] != ws_ref).sum())}, "
              f"top digit range [{got[0].min()}, {got[0].max()}]")
        ok &= bad == 0 and (wsc[o] == ws_ref).all()
        Wd.append(got)

    # ---- K* digits
    TN = 64
    tiles = ((n_cand + TN - 1) // TN + 3) // 4 * 4
    kq = torch.zeros(m * tiles * npad * 6 * TN, dtype=torch.uint8, device="cuda")
    meandot = torch.zeros(m * tiles * TN, dtype=torch.float64, device="cuda")
    cand_dev = to_device(cand)
    _, pv = _lib.host_doubles(var0, m)
    _, pl = _lib.host_doubles(ls, m)
    _lib.check(lib.bo_i8_kstar_digits(kq.data_ptr(), meandot.data_ptr(), cand_dev.data_ptr(), 0, d, n_cand,
                                      gp.x.data_ptr(), gp.x.stride(0), n, d, m, gp.alpha.data_ptr(), pv, pl, None))
    torch.cuda.synchronize()
    kqh = kq.cpu().numpy().reshape(m, tiles, -1)
    Kd = []
    for o in range(m):
        sq = ((x[:n, None, :] - cand[None, :, :]) ** 2).sum(-1)
        kt = np.exp(-0.5 * sq / ls[o] ** 2)  # (n, n_cand)
        got = np.zeros((S, tiles * TN, npad), dtype=np.int64)
        for t in range(tiles):
            got[:, t * TN:(t + 1) * TN, :] = planes_from_buffer(kqh[o, t], TN, nk)
        recon = sum(got[s].astype(np.float64) * 256.0 ** (S - 1 - s) for s in range(S)) * 2.0 ** -46
        err = np.abs(recon[:n_cand, :n].T - kt).max()
        print(f"[K* digits] obj {o}: max |recon - exp| = {err:.3e} (2^-47 = {2.0**-47:.3e}), "
              f"top digit range [{got[0].min()}, {got[0].max()}]")
        ok &= err < 3e-14
        Kd.append(got)

    # ---- the MMA kernel against the exact integer contraction of the digits it was given
    for nsplit in sorted({1, min(2, nb), nb}):
        q_dev = torch.zeros(m * nsplit * tiles * TN, dtype=torch.float64, device="cuda")
        _lib.check(lib.bo_i8_sumsq(q_dev.data_ptr(), gp.wq.data_ptr(), gp.wscale.data_ptr(), kq.data_ptr(), n, m,
                                   n_cand, nsplit, pv, None))
        torch.cuda.synchronize()
        qg = q_dev.cpu().numpy().reshape(m, nsplit, tiles * TN).sum(1)
        for o in range(m):
            acc = [np.zeros((npad, tiles * TN), dtype=np.int64) for _ in range(S)]
            for s in range(S):
                for t in range(S - s):
                    acc[s + t] += Wd[o][s] @ Kd[o][t].T
            b = [(acc[3 * j] * 256 + acc[3 * j + 1]) * 256 + acc[3 * j + 2] for j in range(2)]
            f = wsc[o][:, None]
            v = b[0].astype(np.float64) * f + b[1].astype(np.float64) * (f / 16777216.0)
            want = (v * v).sum(0) * var0[o] ** 2
            rel = np.abs(qg[o] - want).max() / max(np.abs(want).max(), 1e-300)
            print(f"[sumsq nsplit={nsplit}] obj {o}: max rel diff vs exact digit contraction = {rel:.3e}")
            ok &= rel < 1e-13

    # ---- whole pass against the oracle and the DMMA engine
    out = gp.score(cand, betas, want=("mu", "var", "acq"))
    gp2 = DeviceGP(variance_engine="dmma")
    gp2.fit(x, y, mu0, var0, ls, n)
    out2 = gp2.score(cand, betas, want=("mu", "var", "acq"))
    torch.cuda.synchronize()
    want = orc.chol_hot_path(x, y, cand, mu0, var0, ls, betas, n, 3)
    for o in range(m):
        ev = np.abs(out["var"][o].cpu().numpy() - want["var"][o]).max() / var0[o]
        em = np.abs(out["mu"][o].cpu().numpy() - want["mu"][o]).max() / np.sqrt(var0[o])
        e2 = np.abs(out["var"][o].cpu().numpy() - out2["var"][o].cpu().numpy()).max() / var0[o]
        print(f"[score] obj {o}: |var-oracle|/var0 = {ev:.3e}  |mu-oracle|/sd0 = {em:.3e}  |var_i8-var_dmma|/var0 = {e2:.3e}")
        ok &= ev < 1e-9 and em < 1e-9
    print("PROBE", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
