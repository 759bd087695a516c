"""Print the INT8 guard's sampled difference and tolerance for the BASELINE training sets: python tools/guard_tolerances.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200.engine import DeviceGP  # noqa: E402
from bayesopt_smart_b200.workloads import CONFIGS, make_training_set  # noqa: E402

for tag in ("cfg2", "cfg4", "cfg3", "hl"):
    c = CONFIGS[tag]
    x, y, mu0, var0 = make_training_set(c["fn"], c["n"], c["d"], seed=0)
    gp = DeviceGP(variance_engine="int8")
    gp.fit(x, y, mu0, var0, np.full(c["m"], c["ls"]), c["n"])
    cand = torch.rand(400_000, c["d"], dtype=torch.float64, device="cuda")
    gp.score(cand, np.full(c["m"], 2.0), want=("acq",))
    print(json.dumps({"config": tag, "n": c["n"], "cond_measured_at_survey": c["cond"],
                      "parity_tolerance": max(1e-9, 10 * 2.22e-16 * c["cond"]),
                      "guard_tolerance": gp.last_guard_tolerance, "sampled_max_dvar_over_var0": gp.last_guard_worst}),
          flush=True)
