"""Edge cases the reference's own tests exercise (Tests/test_EM.cpp, Tests/test_KMeans.cpp), through the public cppyml
API on the device: one component (mean = data mean, test_EM.cpp:89-98, test_KMeans.cpp:79-88), as many components as
points (exact fit, test_EM.cpp:126-144, test_KMeans.cpp:108-127), assign_responsibilities / assign_label agreeing with
the fitted rows (test_EM.cpp:78-87, test_KMeans.cpp:64-73), plus ragged sizes around the kernels' tile boundaries."""
import numpy as np
import pytest

import oracle
from tests.datasets import synthetic_gmm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clustering():
    from ml_b200 import cabi, import_cppyml
    assert cabi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return import_cppyml().clustering


def test_one_component_is_the_sample_mean_and_covariance(clustering):
    data, _ = oracle.testdata_two_gaussians()
    em = clustering.EM(1)
    em.set_seed(42)
    assert em.fit(data)
    assert np.max(np.abs(em.means[:, 0] - data.mean(axis=0))) <= 1e-14      # test_EM.cpp:96
    assert np.max(np.abs(em.covariance(0) - np.cov(data.T, bias=True))) <= 1e-13
    assert abs(em.mixing_probabilities[0] - 1.0) <= 1e-15
    km = clustering.KMeans(1)
    km.set_seed(42)
    assert km.fit(data)
    assert np.max(np.abs(np.asarray(km.centroids)[0] - data.mean(axis=0))) <= 1e-14   # test_KMeans.cpp:86
    assert set(km.labels) == {0}


def test_as_many_components_as_points_is_an_exact_fit(clustering):
    data = np.ascontiguousarray(np.random.default_rng(3).normal(size=(7, 3)))
    em = clustering.EM(7)
    assert em.fit(data)
    assert np.array_equal(em.means.T, data)
    assert np.array_equal(em.responsibilities, np.eye(7))
    assert np.isinf(em.log_likelihood)
    km = clustering.KMeans(7)
    assert km.fit(data)
    assert np.array_equal(np.asarray(km.centroids), data) and list(km.labels) == list(range(7)) and km.inertia == 0.0
    for model in (clustering.EM(8), clustering.KMeans(8)):
        with pytest.raises(ValueError):
            model.fit(data)                                                   # fewer points than components


# every shape keeps at least ~8 D points per component: below that the covariances are close to singular and the
# reference's own arithmetic is chaotic (see test_em_fixed_steps_match_oracle)
@pytest.mark.parametrize("n,d,k", [(17, 1, 2), (63, 2, 3), (65, 2, 4), (129, 3, 5), (257, 8, 4), (513, 16, 2), (330, 20, 1), (2001, 4, 40)])
def test_ragged_sizes_match_the_oracle(clustering, n, d, k):
    """Sizes around the 16 / 64 / 128-point tiles of the kernels, fused and split paths."""
    data, _, _ = synthetic_gmm(n, d, min(k, 4), seed=n + d, spread=4.0)
    em = clustering.EM(k)
    em.set_seed(5)
    em.set_maximum_steps(4)
    em.fit(data)
    ref = oracle.em_fit(data, k, seed=5, maximum_steps=4)
    if np.all(np.isfinite(ref.means)) and np.isfinite(ref.log_likelihood):
        assert em.number_iterations == ref.iterations
        assert abs(em.log_likelihood - ref.log_likelihood) <= 1e-9 * abs(ref.log_likelihood)
        assert np.max(np.abs(em.means - ref.means)) <= 1e-9 * np.max(np.abs(ref.means))
    km = clustering.KMeans(k)
    km.set_seed(5)
    km.fit(data)
    kref = oracle.kmeans_fit(data, k, seed=5)
    assert km.number_iterations == kref.iterations
    assert np.array_equal(np.asarray(km.labels, dtype=np.uint32), kref.labels)
    assert abs(km.inertia - kref.inertia) <= 1e-12 * max(kref.inertia, 1e-300)


def test_single_point_queries_agree_with_the_fitted_rows(clustering):
    """EM::assign_responsibilities equals the rows of responsibilities() (test_EM.cpp:78-87) once the fit has converged to
    1e-14; KMeans::assign_label reproduces labels and the inertia (test_KMeans.cpp:57-73)."""
    data, _ = oracle.testdata_two_gaussians()
    em = clustering.EM(2)
    em.set_seed(42)
    em.set_absolute_tolerance(1e-15)
    em.set_relative_tolerance(1e-15)
    em.fit(data)
    resp = em.responsibilities
    for i in range(0, 400, 37):
        assert np.max(np.abs(em.assign_responsibilities(data[i]) - resp[i])) <= 1e-12
    km = clustering.KMeans(2)
    km.set_seed(42)
    assert km.fit(data)
    total = 0.0
    for i in range(400):
        label, sq = km.assign_label(data[i])
        assert label == km.labels[i]
        total += sq
    assert abs(total - km.inertia) <= 1e-12 * km.inertia


def test_empty_cluster_goes_to_the_origin():
    """update_step (KMeans.cpp:180-192): a cluster that receives no point ends at the origin."""
    from ml_b200 import cabi
    ctx = cabi.Context(1)
    data = np.ascontiguousarray(np.random.default_rng(8).normal(size=(500, 4)) + 5.0)
    centroids = np.zeros((4, 3))
    centroids[:, 0] = data[0]
    centroids[:, 1] = data[1]
    centroids[:, 2] = 1e6          # nobody is closest to this one
    d_data = cabi.Data.upload(ctx, data)
    km = cabi.Km(d_data, 3)
    km.set_centroids(centroids)
    km.assign()
    km.update()
    assert np.array_equal(km.get_centroids()[:, 2], np.zeros(4))
    ref = oracle.kmeans_fit(data, 3, init=oracle.EXPLICIT, explicit_means=centroids, maximum_steps=2, absolute_tolerance=0.0)
    km.close(); d_data.close(); ctx.close()
    assert ref.iterations == 2
