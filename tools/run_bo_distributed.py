"""Run the drop-in BayesianOptimization loop under torchrun (one process per GPU) and print the batch trace;
used to check that an N-rank run reproduces the 1-rank trace bit for bit (scores are shard-invariant).

    python tools/run_bo_distributed.py                       # single GPU
    torchrun --nproc-per-node 2 tools/run_bo_distributed.py  # candidate set sharded over 2 GPUs
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesopt_smart_b200 as bo  # noqa: E402


def objective(x):
    return np.array([-((x[0] - 40) ** 2) - (x[1] - 25) ** 2 + 0.5 * x[2], -((x[1] - 10) ** 2) - (x[2] - 30) ** 2,
                     -abs(x[0] - x[2]) * 3.0])


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # BO_DIST_BACKEND=gloo lets several ranks share one GPU (NCCL refuses duplicate devices): the 1-GPU test box
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(os.environ.get("BO_DIST_BACKEND", "nccl"))
    trace = []
    np.random.seed(7)  # identical LHS initialisation on every rank
    opt = bo.BayesianOptimization(function=objective, bounds=[(0, 60), (0, 50), (0, 45)], n_objectives=3,
                                  n_iterations=6, initial_samples=12, batch_size=4, betas=[2.0, 2.0, 1.0],
                                  callbacks=[lambda s: trace.append((s["x_next"].tolist(),
                                                                     hashlib.sha1(s["acquisition_values"].tobytes()).hexdigest(),
                                                                     hashlib.sha1(s["mu_objectives"].tobytes()).hexdigest()))])
    opt.optimize()
    rank = dist.get_rank() if world > 1 else 0
    print(json.dumps({"world": world, "rank": rank, "trace": trace,
                      "y_hash": hashlib.sha1(opt.y_vector.tobytes()).hexdigest()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
