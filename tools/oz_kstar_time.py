"""Time the K* digit generator alone (test hook bo_i8_kstar_digits).  BO_LIB overrides the library path.
Usage: python tools/oz_kstar_time.py [n] [n_cand] [d] [m]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesopt_smart_b200 import _lib  # noqa: E402

if os.environ.get("BO_LIB"):
    _lib.LIB_PATH = os.environ["BO_LIB"]
from bayesopt_smart_b200.engine import DeviceGP, to_device  # noqa: E402
from bayesopt_smart_b200 import workloads as orc  # noqa: E402  (input definitions only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 75776
d = int(sys.argv[3]) if len(sys.argv) > 3 else 6
m = int(sys.argv[4]) if len(sys.argv) > 4 else 2
lib = _lib.load()
x, y, mu0, var0 = orc.make_training_set("zdt1" if m == 2 else "dtlz2", n, d, seed=0)
gp = DeviceGP(variance_engine="int8")
gp.fit(x, y, mu0, var0, np.full(m, 0.3), n)
cand = to_device(np.random.default_rng(1).random((n_cand, d)))
npad = lib.bo_npad(n)
tiles = ((n_cand + 63) // 64 + 3) // 4 * 4
kq = torch.empty(m * tiles * npad * 384, dtype=torch.uint8, device="cuda")
md = torch.empty(m * tiles * 64, dtype=torch.float64, device="cuda")
_, pv = _lib.host_doubles(var0, m)
_, pl = _lib.host_doubles(np.full(m, 0.3), m)


def run():
    _lib.check(lib.bo_i8_kstar_digits(kq.data_ptr(), md.data_ptr(), cand.data_ptr(), 0, d, n_cand, gp.x.data_ptr(),
                                      gp.x.stride(0), n, d, m, gp.alpha.data_ptr(), pv, pl, None))


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
us = 1e3 * float(np.median(ts))
print(f"lib={os.path.basename(_lib.LIB_PATH)} n={n} d={d} m={m} n_cand={n_cand}: {us:.1f} us, "
      f"{n_cand * m * npad / us / 1e3:.1f} G entries/s, checksum {int(kq.view(torch.int64)[:4096].sum().item())}")
