"""The EM device path outside the comfortable regime ("weak" item 1): components that sit hundreds to tens of
thousands of their own standard deviations away from the centre of the data (un-standardised features, tight far-apart
clusters), and data with a large common offset.  The reference evaluates (x - mu_k)^T P_k (x - mu_k) directly
(ML/EM.cpp:206-207) and accumulates the covariance about the new mean (EM.cpp:246-248), so it has no cancellation in this
regime; the bar here is the same 1e-9 as everywhere else."""
import numpy as np
import pytest

import oracle
from tests.datasets import separated_clusters
from tests.test_gpu_cabi_parity import em_fit_cabi, rel_err, RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ml_b200 import cabi
    assert cabi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    c = cabi.Context(1)
    yield c
    c.close()


CASES = [(d, sep, off) for d in (2, 8, 16) for sep in (1e2, 1e3, 1e4) for off in (0.0,)] + [(2, 10.0, 1e6), (8, 10.0, 1e6), (16, 1e3, 1e6)]


@pytest.mark.parametrize("d,separation,offset", CASES)
def test_em_far_apart_components_match_oracle(ctx, d, separation, offset):
    k, n = 4, 6000
    data, _, centres = separated_clusters(n, d, k, separation, offset, seed=int(d + separation) % 1000)
    init = np.ascontiguousarray((centres + 0.3).T)
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=60)
    assert np.all(np.isfinite(ref.means)) and np.isfinite(ref.log_likelihood), "the reference itself must be finite on this case"
    fit = em_fit_cabi(ctx, data, k, init, maximum_steps=60)
    # Data with a common offset: the REFERENCE's own arithmetic is conditioned by |x| / sigma (x - mu with both ~1e6 keeps
    # ~1e-10 absolute, and its means are sequential sums of terms ~1e6), so two correct evaluations agree to about
    # |offset| * 1e-14 there, not to 1e-9 of sigma.  The device path works about the data mean and is the more accurate one.
    tol = RTOL * max(1.0, abs(offset) * 2e-5)
    if separation >= 1e3 and offset == 0.0:
        assert 3 in fit.paths, "components this far from the centre of the data must be routed to the direct kernels"
    assert fit.iterations == ref.iterations and fit.converged == ref.converged
    assert abs(fit.log_likelihood - ref.log_likelihood) <= RTOL * abs(ref.log_likelihood)
    assert rel_err(fit.mixing_probabilities, ref.mixing_probabilities) <= RTOL
    for j in range(k):
        # per component, relative to that component's own scale: the means relative to the cluster's standard deviation
        # (not to the huge common offset), the covariances relative to their own largest entry
        sigma = np.sqrt(np.max(np.diag(ref.covariances[j])))
        assert np.max(np.abs(fit.means[:, j] - ref.means[:, j])) <= tol * max(sigma, np.max(np.abs(ref.means[:, j])) * 1e-3), (j, "mean")
        assert rel_err(fit.covariances[j], ref.covariances[j]) <= tol, (j, "covariance")
    assert np.max(np.abs(fit.responsibilities - ref.responsibilities)) <= tol
    if ref.converged:
        assert np.array_equal(fit.labels, ref.labels)
