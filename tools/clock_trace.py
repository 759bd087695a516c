"""Sample SM clock / power every 20 ms while the score pass runs back to back for a few seconds."""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)

x, y, mu0, var0 = orc.make_training_set("zdt1", 1024, 6, seed=0)
gp = DeviceGP()
gp.fit(x, y, mu0, var0, [0.3, 0.3], 1024)
cand = torch.rand(1_000_000, 6, dtype=torch.float64, device="cuda")
out = {"acq": torch.empty(1_000_000, dtype=torch.float64, device="cuda")}
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.active",
                         "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
time.sleep(0.5)
t_start = time.perf_counter()
times = []
for i in range(60):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gp.score(cand, [2.0, 2.0], want=("acq",), out=out)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
t_end = time.perf_counter()
time.sleep(0.3)
proc.terminate()
busy = [l for t, l in lines if t_start <= t <= t_end]
idle = [l for t, l in lines if t < t_start]
clk = [float(l.split(",")[0]) for l in busy]
pw = [float(l.split(",")[1]) for l in busy]
print(json.dumps({"score_ms_first5": times[:5], "score_ms_last5": times[-5:], "score_ms_min": min(times),
                  "samples_busy": len(busy), "clk_busy_min_med_max": [min(clk), float(np.median(clk)), max(clk)],
                  "power_busy_min_med_max": [min(pw), float(np.median(pw)), max(pw)],
                  "idle_sample": idle[-1] if idle else None, "reasons_seen": sorted({l.split(",")[3].strip() for l in busy})}))
