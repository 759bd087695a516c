"""GPU probe: measured FP64 peaks (cuBLAS DGEMM via torch, the library's own DMMA GEMM) and stage timings
of the hot path at the BASELINE shapes.  Writes JSON lines to gpurun_out/probe.jsonl.  Measurement aid only."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesopt_smart_b200 import _lib  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, device_info, to_device  # noqa: E402
from oracle import gp_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "probe.jsonl"), "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    emit(kind="device", **device_info(), name=torch.cuda.get_device_name(0))
    lib = _lib.load()
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty_like(a)
        t = timed(lambda: torch.matmul(a, b, out=c))
        emit(kind="cublas_dgemm", n=n, seconds=t, tflops=2 * n**3 / t / 1e12)
        t = timed(lambda: _lib.check(lib.bo_dgemm_nt_f64(c.data_ptr(), a.data_ptr(), b.data_ptr(), n,
                                                          torch.cuda.current_stream().cuda_stream)))
        emit(kind="own_gemm64", n=n, seconds=t, tflops=2 * n**3 / t / 1e12)
        del a, b, c
    shapes = [("cfg2", "zdt1", 1024, 6, 2, 0.3, 1_000_000), ("cfg4s", "dtlz2", 2048, 8, 3, 0.5, 500_000),
              ("cfg3s", "zdt2", 4096, 10, 2, 0.5, 300_000)]
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if s[0] in sys.argv[1:]]
    for tag, fn, n, d, m, ls, n_cand in shapes:
        x, y, mu0, var0 = orc.make_training_set(fn, n, d, seed=0)
        gp = DeviceGP()
        lsv, betas = np.full(m, ls), np.full(m, 2.0)
        xd, yd = to_device(x), to_device(y)
        t_fit = timed(lambda: gp.fit(xd, yd, mu0, var0, lsv, n), warm=1, reps=3)
        cand = torch.rand(n_cand, d, dtype=torch.float64, device="cuda")
        out = {"acq": torch.empty(n_cand, dtype=torch.float64, device="cuda")}
        t_score = timed(lambda: gp.score(cand, betas, want=("acq",), out=out), warm=1, reps=3)
        flops = n_cand * m * float(n) * n
        t_sel = timed(lambda: gp.topk(out["acq"], 19), warm=1, reps=3)
        emit(kind="hot_path", tag=tag, n=n, d=d, m=m, n_cand=n_cand, fit_s=t_fit, score_s=t_score, topk_s=t_sel,
             cand_per_s=n_cand / (t_score + t_sel), trmm_tflops_equiv=flops / t_score / 1e12)


if __name__ == "__main__":
    main()
