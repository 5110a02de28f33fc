import sys
import numpy as np
sys.path.insert(0, ".")
from ml_b200 import cabi
from tests.datasets import synthetic_gmm

n, d, k, n_sets = [int(a) for a in sys.argv[1:5]]
ctx = cabi.Context(1)
data, _, _ = synthetic_gmm(n, d, max(2, k // 2), seed=n % 97, spread=6.0)
dev = cabi.Data.upload(ctx, data)
rng = np.random.default_rng(d)
starts = [np.ascontiguousarray(data[rng.choice(len(data), size=k, replace=False)].T) for _ in range(n_sets)]
sets = cabi.Kms(dev, k, n_sets)
for s in range(n_sets):
    sets.set_centroids(s, starts[s])
singles = []
for s in range(n_sets):
    km = cabi.Km(dev, k)
    km.set_centroids(starts[s])
    singles.append(km)
for it in range(3):
    inertia, changed = sets.assign()
    for s in range(n_sets):
        i1, c1 = singles[s].assign()
        same_labels = np.array_equal(sets.get_labels(s), singles[s].get_labels())
        print("it", it, "set", s, "inertia", inertia[s] == i1, inertia[s], i1, "changed", changed[s], c1, "labels", same_labels, flush=True)
    shift = sets.update()
    for s in range(n_sets):
        s1 = singles[s].update()
        same_c = np.array_equal(sets.get_centroids(s), singles[s].get_centroids())
        print("it", it, "set", s, "shift", shift[s] == s1, shift[s], s1, "centroids", same_c, float(np.max(np.abs(sets.get_centroids(s) - singles[s].get_centroids()))), flush=True)
print("done")
