"""Time the in-library Pareto filter (bo_pareto_mask_filtered_f64) on 8 M x 3 points: python tools/pareto_time.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200.pareto import pareto_mask_device  # noqa: E402

for kind in ("randn", "sphere_shell"):
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randn(8_000_000, 3, dtype=torch.float64, device="cuda", generator=g)
    if kind == "sphere_shell":  # a much larger front: points near a sphere octant
        y = y.abs()
        y = y / y.norm(dim=1, keepdim=True) * (1.0 - 0.05 * torch.rand(8_000_000, 1, dtype=torch.float64, device="cuda", generator=g))
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mask = pareto_mask_device(y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps({"kind": kind, "n": 8_000_000, "m": 3, "ms": min(ts[1:]), "front": int(mask.sum().item())}), flush=True)
