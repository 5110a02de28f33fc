"""The reference's own K-means benchmark (Benchmarks/bm_KMeans.cpp:9-47): mouse data, D = 2, K = 3, K-means++ start, tolerance
1e-14, THREE initialisations per fit, N = 100 .. 100000 (and beyond).  Times `cppyml.clustering.KMeans.fit` with the starts in
lockstep (one pass of the assignment kernel per iteration for all of them, mlb_kms) and one after the other
(MLPP_KMEANS_LOCKSTEP=0), next to the oracle port on one host core.  One JSON line per N."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fit_seconds(n, repeats):
    import oracle
    from ml_b200 import import_cppyml
    cppyml = import_cppyml()
    data, _ = oracle.testdata_mouse(n)
    out = []
    for rep in range(repeats + 1):
        km = cppyml.clustering.KMeans(3)
        km.set_seed(42)
        km.set_absolute_tolerance(1e-14)
        km.set_centroids_initialiser(cppyml.clustering.KPP())
        km.set_number_initialisations(3)
        t0 = time.perf_counter()
        km.fit(data)
        out.append(time.perf_counter() - t0)
    return min(out[1:]), km.inertia, km.number_iterations


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        n = int(sys.argv[2])
        print(json.dumps(fit_seconds(n, 5)))
        sys.exit(0)
    import oracle
    for n in [int(a) for a in sys.argv[1:]] or [100, 1000, 10000, 100000, 1000000]:
        row = {"n": n, "d": 2, "k": 3, "initialisations": 3}
        for name, flag in (("lockstep", "1"), ("one_after_the_other", "0")):
            env = dict(os.environ, MLPP_KMEANS_LOCKSTEP=flag)
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n)], env=env, capture_output=True, text=True, check=True)
            sec, inertia, iters = json.loads(res.stdout.strip().splitlines()[-1])
            row[name + "_ms"] = sec * 1e3
            row[name + "_inertia"] = inertia
        data, _ = oracle.testdata_mouse(n)
        t0 = time.perf_counter()
        ref = oracle.kmeans_fit(data, 3, init=oracle.KPP, seed=42, absolute_tolerance=1e-14, number_initialisations=3)
        row["oracle_one_core_ms"] = (time.perf_counter() - t0) * 1e3
        row["oracle_inertia"] = ref.inertia
        row["same_inertia"] = bool(abs(row["lockstep_inertia"] - ref.inertia) <= 1e-12 * ref.inertia and row["lockstep_inertia"] == row["one_after_the_other_inertia"])
        print(json.dumps(row), flush=True)
