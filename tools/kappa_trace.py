"""kappa_max (the routing statistic of the EM step, em.cu kKappaDirect) along bench.py's own C3 trajectory: data points
0..K-1 as initial means, sample covariance, N in argv.  One JSON line per N."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_b200 import cabi  # noqa: E402

D, K, STEPS = 16, 32, int(os.environ.get("STEPS", "40"))
ctx = cabi.Context(1)
for n in [int(a) for a in sys.argv[1:]]:
    data = cabi.Data.generate_gmm(ctx, n, D, 32, seed=20261018)
    em = cabi.Em(data, K)
    cov = em.sample_covariance()
    em.set_params(np.ascontiguousarray(data.download(0, K).T), np.repeat(cov[None], K, axis=0), np.full(K, 1.0 / K))
    em.set_kernel_timing(True)
    rows = []
    prev_ms, prev_n = 0.0, 0
    for step in range(STEPS):
        ll = em.step()
        kappa, nxt = em.conditioning()
        ms, launches = em.kernel_time_ms()
        rows.append({"step": step, "ll": ll, "path": em.last_path, "kappa_after": kappa, "next": nxt, "kernel_ms": ms - prev_ms})
        prev_ms = ms
    _, _, weights = em.get_params()
    print(json.dumps({"n": n, "min_weight": float(weights.min()), "trace": rows}), flush=True)
    em.close(); data.close()
