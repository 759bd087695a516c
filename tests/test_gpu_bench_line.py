"""The driver's bench contract on a real device: `python bench.py` prints one JSON line with the required keys,
the roofline / cpu_baseline / e2e objects, a non-zero launch count and the second (INT8 engine) measurement."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "int8_engine"):
        assert key in line, key
    assert line["unit"] == "candidates/s" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["scaling"] == "weak"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] > 0 and line["value"] > 1e6
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= 1.02 * line["value"]
    roof = line["roofline"]
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and roof["kernel"] == "trmm_sumsq_kernel"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and 0.5 < roof["frac"] < 1.1
    assert roof["traffic"] is None or roof["traffic"] > 0  # flat: DRAM bytes per launch from the ncu capture
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0 and "sample" in cpu
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "bayesopt")):
        assert cpu["kind"] == "reference", cpu.get("why_port")  # the reference's own Numba code ran on this box
    # the other BASELINE configs ride in the same line, each with its two correctness assertions
    extras = line["baseline_configs"]
    for label in ("north_star", "cfg4"):
        for engine in ("dmma", "int8"):
            blk = extras[label]["engines"][engine]
            assert "error" not in blk, blk
            assert blk["sharded_topk_equals_gathered"] is True
            assert blk["oracle_spot_check"]["within_tolerance"] is True
            assert blk["scaling"] == "strong" and blk["value"] > 0
        assert extras[label]["engines"]["dmma"]["batch_idx"] == extras[label]["engines"]["int8"]["batch_idx"]
    assert extras["north_star"]["engines"]["dmma"]["frac_of_fp64_ceiling"] > 0.85
    assert extras["cfg5"]["all_finite"] is True and extras["cfg5"]["roofline"]["frac"] > 0.5
    assert extras["hbm_passes"]["score_pass_m2_16M"]["frac"] > 0.7
    assert extras["cfg1_loop"]["pareto_front"] == [[100.0, 20.0]]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    i8 = line["int8_engine"]
    assert i8["same_top_batch_as_dmma"] is True and i8["max_abs_acq_difference_vs_dmma"] < 1e-8
    assert i8["value"] > line["value"] and i8["gpu_launches"] > 0
    assert i8["roofline"]["kernel"] == "oz_sumsq_kernel" and 0.2 < i8["roofline"]["frac"] < 1.1
    assert i8["e2e"]["h2d_bytes_per_step"] == e2e["h2d_bytes_per_step"]
