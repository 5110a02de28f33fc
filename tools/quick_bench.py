"""Scratch timing of the device paths on generated data (not the bench contract; see bench.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ml_b200 import cabi

F_EM = lambda d: 2 * d * d + 8 * d + 6


def em_case(ctx, n, d, k, steps=10):
    data = cabi.Data.generate_gmm(ctx, n, d, k, seed=7)
    em = cabi.Em(data, k)
    cov = em.sample_covariance()
    init = data.download(0, k).T
    em.set_params(init, np.repeat(cov[None], k, axis=0), np.full(k, 1.0 / k))
    em.run_steps(3)
    ctx.timer_start()
    lls = em.run_steps(steps, want_ll=True)
    ms = ctx.timer_stop() / steps
    out = dict(kind="em", n=n, d=d, k=k, ms_per_iter=ms, gpc_per_s=n * k / ms / 1e6, tflops=n * k * F_EM(d) / ms / 1e9,
               hbm_gbs=8 * d * n / ms / 1e6, ll_last=float(lls[-1]))
    em.close(); data.close()
    return out


def km_case(ctx, n, d, k, steps=5):
    data = cabi.Data.generate_gmm(ctx, n, d, min(k, 64), seed=9)
    km = cabi.Km(data, k)
    km.set_centroids(data.download(0, k).T)
    for _ in range(2):
        km.assign(); km.update()
    ctx.timer_start()
    for _ in range(steps):
        inertia, changed = km.assign(); km.update()
    ms = ctx.timer_stop() / steps
    out = dict(kind="kmeans", n=n, d=d, k=k, ms_per_iter=ms, gpc_per_s=n * k / ms / 1e6, tflops=n * k * (3 * d + 1) / ms / 1e9,
               hbm_gbs=(8 * d + 4) * n / ms / 1e6, inertia=inertia, changed=changed)
    km.close(); data.close()
    return out


if __name__ == "__main__":
    ctx = cabi.Context(1)
    cases = sys.argv[1:] or ["em:10000000:8:16", "em:10000000:16:32", "km:10000000:32:256"]
    for c in cases:
        kind, n, d, k = c.split(":")
        t0 = time.time()
        r = (em_case if kind == "em" else km_case)(ctx, int(n), int(d), int(k))
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)
