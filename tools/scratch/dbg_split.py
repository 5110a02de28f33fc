import sys, numpy as np
sys.path.insert(0, ".")
import oracle
from ml_b200 import cabi
from tests.datasets import synthetic_gmm
ctx = cabi.Context(1)
for (n, d, k) in [(6000, 64, 34), (6000, 64, 4), (6000, 48, 34), (6000, 56, 8), (6000, 60, 8), (6000, 64, 8), (20000, 64, 34)]:
    data, _, _ = synthetic_gmm(n, d, k, seed=35, spread=6.0)
    init = np.ascontiguousarray(data[:: n // k][:k].T)
    steps = 2
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, absolute_tolerance=0.0, relative_tolerance=0.0, want_responsibilities=False)
    dd = cabi.Data.upload(ctx, data)
    em = cabi.Em(dd, k)
    cov = em.sample_covariance()
    em.set_params(init, np.repeat(cov[None], k, axis=0), np.full(k, 1.0 / k))
    lls = [em.step() for _ in range(steps)]
    m, c, w = em.get_params()
    print(n, d, k, "ll", lls, "ref", ref.log_likelihood, "means err", np.abs(m - ref.means).max(), "w err", np.abs(w - ref.mixing_probabilities).max(), flush=True)
    em.close(); dd.close()
