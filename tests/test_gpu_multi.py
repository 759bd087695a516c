"""Multi-GPU shard invariance through the public API: the BO loop under torchrun with 2 ranks must reproduce
the single-process trace bit for bit (same batches, same per-candidate arrays, same observations).
Needs >= 2 GPUs on the box; skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tools", "run_bo_distributed.py")


def _last_json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


@pytest.mark.timeout(600)
def test_two_rank_bo_loop_equals_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    one = subprocess.run([sys.executable, SCRIPT], capture_output=True, text=True, env=env, timeout=300)
    assert one.returncode == 0, one.stderr[-2000:]
    ref = _last_json_lines(one.stdout)[-1]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", SCRIPT], capture_output=True,
                         text=True, env=env, timeout=300)
    assert two.returncode == 0, two.stderr[-2000:]
    rows = _last_json_lines(two.stdout)
    assert sorted(r["rank"] for r in rows) == [0, 1]
    for r in rows:
        assert r["trace"] == ref["trace"]
        assert r["y_hash"] == ref["y_hash"]
