"""Achieved HBM bandwidth of the memory-bound passes (stand-alone score pass a6..a8, top-k scan, Pareto prefilter)
against the measured copy bandwidth in MEASURED_PEAKS.json.  Writes gpurun_out/hbm_passes.jsonl."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesopt_smart_b200 import _lib  # noqa: E402
from bayesopt_smart_b200.engine import DeviceGP, _ptr, _stream  # noqa: E402
from bayesopt_smart_b200.pareto import _mask_against  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "hbm_passes.jsonl"), "a")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    PEAK_SRC = "MEASURED_PEAKS.json"
except Exception:  # noqa: BLE001
    PEAK, PEAK_SRC = 6650.0, "fallback (B200_PROFILING.md)"


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


def timed(fn, flush, warm=2, reps=7):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        flush.zero_()  # evict L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    lib = _lib.load()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    gp = DeviceGP()
    big = torch.empty(2 << 30, dtype=torch.uint8, device="cuda").view(torch.float64)
    t = timed(lambda: big.fill_(1.0), flush)
    emit(kind="write_only_fill_2GiB", bytes=big.numel() * 8, seconds=t, gbs=big.numel() * 8 / t / 1e9, peak_gbs=PEAK,
         frac=big.numel() * 8 / t / 1e9 / PEAK)
    src = torch.empty_like(big)
    t = timed(lambda: big.copy_(src), flush)
    emit(kind="copy_2GiB_read_plus_write", bytes=2 * big.numel() * 8, seconds=t, gbs=2 * big.numel() * 8 / t / 1e9,
         peak_gbs=PEAK, frac=2 * big.numel() * 8 / t / 1e9 / PEAK)
    t = timed(lambda: big.sum(), flush)
    emit(kind="read_only_sum_2GiB", bytes=big.numel() * 8, seconds=t, gbs=big.numel() * 8 / t / 1e9, peak_gbs=PEAK,
         frac=big.numel() * 8 / t / 1e9 / PEAK)
    del big, src
    for m, M in [(2, 16_000_000), (3, 8_000_000)]:
        mu = torch.randn(m, M, dtype=torch.float64, device="cuda")
        var = torch.rand(m, M, dtype=torch.float64, device="cuda")
        smu, svar, ucb = torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(mu)
        acq = torch.empty(M, dtype=torch.float64, device="cuda")
        _, pm = _lib.host_doubles(np.zeros(m), m)
        _, pv = _lib.host_doubles(np.ones(m) * 2.0, m)
        _, pb = _lib.host_doubles(np.ones(m) * 2.0, m)

        def full():
            _lib.check(lib.bo_acquisition_f64(_ptr(smu), _ptr(svar), _ptr(ucb), _ptr(acq), _ptr(mu), _ptr(var), M, M,
                                              m, pm, pv, pb, _stream()))

        def acq_only():
            _lib.check(lib.bo_acquisition_f64(None, None, None, _ptr(acq), _ptr(mu), _ptr(var), M, M, m, pm, pv, pb,
                                              _stream()))

        t = timed(full, flush)
        b = (2 * m + 3 * m + 1) * 8 * M
        emit(kind="score_pass_all_outputs", m=m, n_cand=M, bytes=b, seconds=t, gbs=b / t / 1e9, peak_gbs=PEAK,
             frac=b / t / 1e9 / PEAK, peak_source=PEAK_SRC, cand_per_s=M / t)
        t = timed(acq_only, flush)
        b = (2 * m + 1) * 8 * M
        emit(kind="score_pass_acq_only", m=m, n_cand=M, bytes=b, seconds=t, gbs=b / t / 1e9, peak_gbs=PEAK,
             frac=b / t / 1e9 / PEAK, cand_per_s=M / t)
        for k in (3, 19):
            t = timed(lambda: gp.topk(acq, k), flush)
            b = 8 * M
            emit(kind="topk_scan", k=k, n_cand=M, bytes=b, seconds=t, gbs=b / t / 1e9, peak_gbs=PEAK,
                 frac=b / t / 1e9 / PEAK)
        rows = ucb.T.contiguous()
        front = rows[:4096][torch.randperm(4096, device="cuda")[:128]]
        front = front[torch.argsort(front.sum(dim=1), descending=True)].contiguous()  # as pareto_mask_device does
        t = timed(lambda: _mask_against(rows, front), flush)
        b = (8 * m + 1) * M
        emit(kind="pareto_prefilter_vs_128_points", m=m, n=M, bytes=b, seconds=t, gbs=b / t / 1e9, peak_gbs=PEAK,
             frac=b / t / 1e9 / PEAK)
        del mu, var, smu, svar, ucb, acq, rows


if __name__ == "__main__":
    main()
