"""BASELINE config 5 as it is worded: ml::Clustering::KMeans Lloyd iterations through the cppyml Python binding (one process;
MLPP_CUDA_DEVICES = number of GPUs the process shards over, default 1).  The data is a numpy float64 C-contiguous (N, D)
array, as cppyml/clustering.cpp:149-184 takes it; the whole fit() call is timed (upload from pageable numpy memory,
Forgy initialisation on the host, `steps` Lloyd iterations, labels and centroids back).

  MLPP_CUDA_DEVICES=1 python tools/c5_cppyml.py [points_per_gpu] [steps]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ml_b200 import cabi, import_cppyml

gpus = int(os.environ.get("MLPP_CUDA_DEVICES", "1"))
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000) * gpus
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d, k = 32, 256

# the bench generator's mixture, brought to the host once (not timed)
ctx = cabi.Context(1)
gen = cabi.Data.generate_gmm(ctx, n, d, 64, seed=20261018)
data = np.empty((n, d))
block = 1 << 22
for lo in range(0, n, block):
    data[lo:lo + block] = gen.download(lo, min(block, n - lo))
gen.close(); ctx.close()

clustering = import_cppyml().clustering
out = {"workload": f"c5 via cppyml: KMeans({k}).fit on a numpy (N={n}, D={d}) array, Forgy initialiser, seed 1, {steps} Lloyd iterations", "gpus": gpus}
for attempt in ("warm-up", "timed"):
    km = clustering.KMeans(k)
    km.set_seed(1)
    km.set_maximum_steps(steps)
    km.set_absolute_tolerance(0.0)
    t0 = time.perf_counter()
    km.fit(data)
    seconds = time.perf_counter() - t0
    out[attempt] = {"fit_seconds": seconds, "iterations": km.number_iterations, "gpc_per_s_e2e": n * k * km.number_iterations / seconds / 1e9,
                    "inertia": km.inertia}
    t0 = time.perf_counter()
    labels, dist = km.assign_labels(data[:4_000_000])
    out[attempt]["assign_labels_4M_seconds"] = time.perf_counter() - t0
    # (a fit that stops at maximum_steps leaves labels of the LAST assignment and centroids of the update after it,
    # KMeans.cpp:80-109, so the batch is checked against the single-point API, not against km.labels)
    for i in range(0, 4_000_000, 400_000):
        label, sq = km.assign_label(data[i])
        assert label == labels[i] and abs(sq - dist[i]) <= 1e-12 * sq
print(json.dumps(out))
