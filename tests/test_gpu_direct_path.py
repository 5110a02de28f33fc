"""The direct-difference EM kernels (ml_b200/csrc/em_direct.cuh: per-component centring, as ML/EM.cpp:205-207 and :246-248)
against the CPU oracle: forced on the shapes the feature-space kernels also take, and on the shapes only they take
(D > 64, K > 256).  Same bar as everywhere: identical iteration count and labels, parameters within 1e-9."""
import numpy as np
import pytest

import oracle
from tests.datasets import synthetic_gmm
from tests.test_gpu_cabi_parity import EM_SHAPES, EM_SPLIT_SHAPES, RTOL, em_fit_cabi, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ml_b200 import cabi
    assert cabi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    c = cabi.Context(1)
    yield c
    c.close()


def check_against(ref, got, k, labels=True):
    assert got.iterations == ref.iterations and got.converged == ref.converged
    assert abs(got.log_likelihood - ref.log_likelihood) <= RTOL * abs(ref.log_likelihood)
    assert rel_err(got.means, ref.means) <= RTOL
    assert rel_err(got.mixing_probabilities, ref.mixing_probabilities) <= RTOL
    for c in range(k):
        assert rel_err(got.covariances[c], ref.covariances[c]) <= RTOL, c
    if labels and ref.converged:
        assert np.array_equal(got.labels, ref.labels)


@pytest.mark.parametrize("n,d,k,seed", EM_SHAPES + EM_SPLIT_SHAPES)
def test_forced_direct_fixed_steps_match_oracle(ctx, n, d, k, seed):
    data, _, true_means = synthetic_gmm(n, d, k, seed=seed, spread=6.0)
    init = np.ascontiguousarray(data[:: n // k][:k].T)
    if d >= 48:
        init = np.ascontiguousarray(true_means.T)   # see test_em_fixed_steps_match_oracle
    steps = 6 if d * d * k <= 20000 else 3
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, absolute_tolerance=0.0, relative_tolerance=0.0)
    got = em_fit_cabi(ctx, data, k, init, absolute_tolerance=0.0, relative_tolerance=0.0, maximum_steps=steps, force_path=3)
    assert set(got.paths) == {3}
    assert rel_err(got.sample_covariance, np.cov(data.T).reshape(d, d)) <= 1e-12
    check_against(ref, got, k, labels=False)
    assert np.max(np.abs(got.responsibilities - ref.responsibilities)) <= 1e-9


@pytest.mark.parametrize("n,d,k,seed", [(10000, 2, 3, 21), (30000, 16, 32, 23), (12000, 24, 12, 24)])
def test_forced_direct_full_fit_matches_oracle(ctx, n, d, k, seed):
    data, _, _ = synthetic_gmm(n, d, k, seed=seed, spread=8.0)
    init = np.ascontiguousarray(data[7:: n // k][:k].T)
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=300)
    got = em_fit_cabi(ctx, data, k, init, maximum_steps=300, force_path=3)
    check_against(ref, got, k)
    assert np.max(np.abs(got.responsibilities - ref.responsibilities)) <= 1e-9
    assert np.max(np.abs(got.responsibilities.sum(axis=1) - 1.0)) <= 1e-14


# shapes the feature-space kernels do not take: D > 64, K > 256
@pytest.mark.parametrize("n,d,k,seed,steps", [(4000, 96, 8, 41, 3), (12000, 16, 600, 42, 3), (3000, 128, 4, 43, 3), (12000, 66, 20, 44, 2), (2000, 65, 1, 45, 3)])
def test_large_shapes_match_oracle(ctx, n, d, k, seed, steps):
    data, _, true_means = synthetic_gmm(n, d, min(k, 40), seed=seed, spread=6.0)
    if k <= 40:
        init = np.ascontiguousarray(true_means.T)
    else:
        init = np.ascontiguousarray(data[:: n // k][:k].T)
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, absolute_tolerance=0.0, relative_tolerance=0.0)
    got = em_fit_cabi(ctx, data, k, init, absolute_tolerance=0.0, relative_tolerance=0.0, maximum_steps=steps)
    assert set(got.paths) == {3}
    assert rel_err(got.sample_covariance, np.cov(data.T).reshape(d, d)) <= 1e-12
    check_against(ref, got, k, labels=False)
    assert np.max(np.abs(got.responsibilities - ref.responsibilities)) <= 1e-9


def test_dimension_limit_is_reported(ctx):
    from ml_b200 import cabi
    data = np.zeros((300, 129))
    d_data = cabi.Data.upload(ctx, data)
    with pytest.raises(cabi.MlbError) as e:
        cabi.Em(d_data, 2)
    assert e.value.code == cabi.MLB_EINVAL and "D <= 128" in str(e.value)
    d_data.close()


@pytest.mark.parametrize("n,d,k,force", [(5000, 8, 16, 3), (3000, 80, 5, 0), (30000, 4, 260, 0)])
def test_direct_maximise_first_and_predict(ctx, n, d, k, force):
    """The `maximise_first` start (EM.cpp:120-125: ClosestCentroid responsibilities, then an M-step) and batched
    prediction (EM.cpp:176-188) on the direct kernels."""
    from ml_b200 import cabi
    data, _, _ = synthetic_gmm(n, d, min(k, 32), seed=5 + d, spread=6.0)
    centroids = np.ascontiguousarray(data[3:: n // k][:k])            # (K, D)
    d2 = ((data[:, None, :] - centroids[None, :, :]) ** 2).sum(axis=2) if n * k * d < 5e7 else np.stack([((data - c) ** 2).sum(axis=1) for c in centroids], axis=1)
    nearest = np.argmin(d2, axis=1)
    resp0 = np.zeros((n, k))
    resp0[np.arange(n), nearest] = 1.0
    ref = oracle.em_fit(data, k, maximise_first=True, resp_init_centroids=oracle.EXPLICIT, explicit_means=centroids.T, maximum_steps=3,
                        absolute_tolerance=0.0, relative_tolerance=0.0)
    d_data = cabi.Data.upload(ctx, data)
    em = cabi.Em(d_data, k)
    if force:
        em.force_path(force)
    em.mstep_from_responsibilities(resp0)
    for _ in range(3):
        ll = em.step()
    assert em.last_path == 3
    means, covs, weights = em.get_params()
    assert abs(ll - ref.log_likelihood) <= RTOL * abs(ref.log_likelihood)
    assert rel_err(means, ref.means) <= RTOL
    assert rel_err(weights, ref.mixing_probabilities) <= RTOL
    for c in range(k):
        assert rel_err(covs[c], ref.covariances[c]) <= RTOL, c
    resp, lab = em.emit()
    assert np.max(np.abs(resp - ref.responsibilities)) <= 1e-9
    queries = data[:257]
    want = np.array([oracle.em_assign_responsibilities(ref, q) for q in queries])
    got, got_labels = em.predict(queries)
    assert np.max(np.abs(got - want)) <= 1e-9
    assert np.array_equal(got_labels, np.argmax(want, axis=1))
    em.close()
    d_data.close()


# ---------------------------------------------------------------- K-means with wide points (D > 64): the reference's scan

@pytest.mark.parametrize("n,d,k,seed", [(6000, 128, 64, 81), (5000, 65, 3, 82), (4000, 100, 400, 83)])
def test_kmeans_wide_points_match_oracle(ctx, n, d, k, seed):
    """K-means beyond D = 64 (km_assign_exact_kernel; K = 400 at D = 100 also needs statistics blocks)."""
    from tests.test_gpu_cabi_parity import _kmeans_blocked_case
    _kmeans_blocked_case(ctx, n, d, k, seed)


def test_kmeans_dimension_limit_is_reported(ctx):
    from ml_b200 import cabi
    d_data = cabi.Data.upload(ctx, np.zeros((300, 129)))
    with pytest.raises(cabi.MlbError) as e:
        cabi.Km(d_data, 2)
    assert e.value.code == cabi.MLB_EINVAL and "D <= 128" in str(e.value)
    d_data.close()
