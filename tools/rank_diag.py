"""Diagnosis of a slow multi-rank step (r02i/r02j: the fused C3 kernel took 35 ms per 12.5M points under torchrun with 2 and 8
ranks, 8.4 ms on one GPU).  Run under torchrun with 2 ranks: times the same 12.5M-point C3 share per rank in several set-ups
and prints one JSON line per set-up and rank."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_b200 import cabi  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
use_torch = os.environ.get("DIAG_TORCH", "1") == "1"
N_PER, D, K = 12_500_000, 16, 32


def exchange_uid():
    path = "/tmp/mlb_diag_uid_%s" % os.environ.get("MASTER_PORT", "0")
    if rank == 0:
        uid = cabi.nccl_unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(bytes(uid))
        os.replace(path + ".tmp", path)
        return uid
    while not os.path.exists(path):
        time.sleep(0.01)
    return open(path, "rb").read()


def measure(tag, ctx, n_total, steps=6, timing=True):
    data = cabi.Data.generate_gmm(ctx, n_total, D, 32, seed=20261018)
    _, n_local, _ = data.shape
    em = cabi.Em(data, K)
    cov = em.sample_covariance()
    # the same fixed start on every rank
    init = np.ascontiguousarray(np.linspace(-8, 8, K * D).reshape(K, D).T)
    em.set_params(init, np.repeat(cov[None], K, axis=0), np.full(K, 1.0 / K))
    em.run_steps(3)
    em.set_kernel_timing(timing)
    ctx.timer_start()
    em.run_steps(steps)
    total = ctx.timer_stop()
    kms, launches = em.kernel_time_ms() if timing else (0.0, 1)
    kappa, path = em.conditioning()
    last = em.last_path
    last = last() if callable(last) else last
    print(json.dumps({"tag": tag, "rank": rank, "n_local": int(n_local), "ms_per_step": total / steps, "kernel_ms_avg": kms / max(1, launches),
                      "last_path": last, "next_path": path, "kappa_max": kappa}), flush=True)
    em.close(); data.close()


if use_torch:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [cabi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
else:
    uid = exchange_uid()

# 1. each rank alone on its GPU (world = 1 context), the other rank doing the same at the same time
ctx = cabi.Context.for_rank(local, 0, 1, None)
measure("alone_world1", ctx, N_PER)
ctx.close()
# 2. the job context
ctx = cabi.Context.for_rank(local, rank, world, uid)
measure("job_timed", ctx, N_PER * world)
measure("job_untimed", ctx, N_PER * world, timing=False)
ctx.close()
# 3. alone again, after NCCL has been up in this process
ctx = cabi.Context.for_rank(local, 0, 1, None)
measure("alone_after", ctx, N_PER)
ctx.close()
if use_torch:
    dist.barrier()
    dist.destroy_process_group()
