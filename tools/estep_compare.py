"""E-step only, the library's DMMA formulation: the emit kernel (em_kernel<16,32,2>: log-densities, log-sum-exp, labels) on
C3's per-GPU share.  Run under `ncu --metrics gpu__time_duration.sum` and read the em_kernel<..., 2> lines; the scalar
formulation is tools/estep_dfma.cu.  Usage: python tools/estep_compare.py [points]"""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_b200 import cabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
ctx = cabi.Context(1)
data = cabi.Data.generate_gmm(ctx, n, 16, 32, seed=20261018)
em = cabi.Em(data, 32)
cov = em.sample_covariance()
em.set_params(np.ascontiguousarray(data.download(0, 32).T), np.repeat(cov[None], 32, axis=0), np.full(32, 1 / 32))
em.run_steps(3)
_, labels = em.emit(want_responsibilities=False, want_labels=True)   # E-step only, staged 1M points at a time
print("labels", labels[:8], "stages of 2^20 points:", (n + (1 << 20) - 1) >> 20)
