"""The fused UCB + exact-HVI pass on its own (16 M candidates, 2 objectives, 300-point front), for timing / ncu:
    python tools/hvi_pass.py [n_cand] [front_points] [m]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200 import acquisition as aq  # noqa: E402
from bayesopt_smart_b200.engine import HviFront  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 300
m = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(0)
if m == 2:
    t = np.sort(rng.random(P))
    pts = np.stack([t, 1.0 - t ** 2], axis=1) * 3.0 - 1.0
else:
    a, b = rng.random(P) * np.pi / 2, rng.random(P) * np.pi / 2
    pts = np.stack([np.cos(a) * np.cos(b), np.sin(a) * np.cos(b), np.sin(b)], axis=1) * 3.0 - 1.0
front = HviFront(pts, np.full(m, -1.5))
mu = torch.randn(m, M, dtype=torch.float64, device="cuda")
var = torch.rand(m, M, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
best = 1e30
for i in range(4):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ucb, hvi = aq.ucb_and_exact_hvi_device(mu, var, np.zeros(m), np.full(m, 2.0), np.full(m, 2.0), front)
    e1.record()
    torch.cuda.synchronize()
    if i:
        best = min(best, e0.elapsed_time(e1) * 1e-3)
nbytes = M * 8 * (2 * m + m + 1)  # mu, var read; ucb, hvi written
print(json.dumps({"kind": "fused_ucb_exact_hvi", "m": m, "n_cand": M, "front_points": int(front.count.item()),
                  "seconds": best, "gbs": nbytes / best / 1e9, "cand_per_s": M / best,
                  "positive_fraction": float((hvi > 0).double().mean().item())}))
