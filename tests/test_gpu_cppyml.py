"""The reference's own clustering tests, restated on the public API of this build (cppyml over the C++ host
classes over the C-ABI over the CUDA kernels), plus seed-for-seed parity with the CPU oracle.

Restated: Tests/test_EM.cpp:8-124, Tests/test_KMeans.cpp:8-106, cppyml/tests/test_clustering.py:47-95.
"""
import numpy as np
import pytest

import oracle
from tests.datasets import mouse_numpy, synthetic_gmm

pytestmark = pytest.mark.gpu

MEANS_TRUE = np.array([[0.4, -1.2], [0.11, 2.2], [0.5, 1.6]])
SIGMAS_TRUE = np.array([[0.05, 0.2], [0.04, 0.1], [0.01, 0.2]])
P0 = 0.25


@pytest.fixture(scope="module")
def clustering():
    from ml_b200 import cabi, import_cppyml
    assert cabi.device_count() >= 1
    return import_cppyml().clustering


def initialiser(clustering, kind):
    return {oracle.FORGY: clustering.Forgy, oracle.RANDOM_PARTITION: clustering.RandomPartition, oracle.KPP: clustering.KPP}[kind]()


@pytest.mark.parametrize("kind,maximise_first", [(oracle.FORGY, False), (oracle.RANDOM_PARTITION, False), (oracle.KPP, False), (None, True)])
def test_em_two_gaussians(clustering, kind, maximise_first):
    """Tests/test_EM.cpp:8-104."""
    data, _ = oracle.testdata_two_gaussians()
    em = clustering.EM(2)
    em.set_absolute_tolerance(1e-8)
    em.set_relative_tolerance(1e-8)
    em.set_maximum_steps(100)
    em.set_seed(63413131)
    if kind is not None:
        em.set_means_initialiser(initialiser(clustering, kind))
    em.set_maximise_first(maximise_first)
    assert em.fit(data)
    assert em.mixing_probabilities.shape == (2,) and em.means.shape == (3, 2) and em.responsibilities.shape == (400, 2)
    resp = em.responsibilities
    for i in range(400):
        assert np.linalg.norm(em.assign_responsibilities(data[i]) - resp[i]) <= 1e-15, i
    means, mix = MEANS_TRUE.copy(), np.array([P0, 1 - P0])
    covs = [np.diag(SIGMAS_TRUE[:, k] ** 2) for k in range(2)]
    if (em.mixing_probabilities[0] < em.mixing_probabilities[1]) != (P0 < 1 - P0):
        mix, means, covs = mix[::-1], means[:, ::-1], covs[::-1]
    assert np.linalg.norm(mix - em.mixing_probabilities) <= 2e-2
    assert np.linalg.norm(means - em.means) <= 2e-2
    for k in range(2):
        assert np.linalg.norm(covs[k] - em.covariance(k)) <= 1e-2
    # seed-for-seed against the oracle: same pseudo-random start, so the same fit
    ref = oracle.em_fit(data, 2, seed=63413131, maximum_steps=100, maximise_first=maximise_first,
                        means_init=oracle.FORGY if kind is None else kind)
    assert em.number_iterations == ref.iterations
    assert abs(em.log_likelihood - ref.log_likelihood) <= 1e-9 * abs(ref.log_likelihood)
    assert np.max(np.abs(em.means - ref.means)) <= 1e-9
    # one component: the mean is the data mean
    em1 = clustering.EM(1)
    if kind is not None:
        em1.set_means_initialiser(initialiser(clustering, kind))
    em1.set_maximise_first(maximise_first)
    em1.fit(data)
    assert em1.log_likelihood <= em.log_likelihood
    assert np.linalg.norm(data.mean(axis=0) - em1.means[:, 0]) <= 1e-14
    r1 = em1.responsibilities
    for i in range(400):
        assert np.linalg.norm(em1.assign_responsibilities(data[i]) - r1[i]) <= 1e-15


def test_em_mouse_matches_sklearn(clustering):
    """cppyml/tests/test_clustering.py:47-74."""
    from sklearn.mixture import GaussianMixture
    data = mouse_numpy()
    em = clustering.EM(3)
    em.set_seed(42)
    em.set_means_initialiser(clustering.KPP())
    em.set_absolute_tolerance(1e-10)
    em.set_relative_tolerance(0)
    em.set_maximum_steps(1000)
    assert em.fit(data)
    gm = GaussianMixture(n_components=3, tol=1e-10, reg_covar=1e-15, random_state=999, max_iter=1000)
    gm.fit(data)
    assert abs(em.log_likelihood - gm.score(data)) <= 1e-10
    u = em.assign_responsibilities(np.array([0.0, 0.0]))
    assert abs(u.sum() - 1) <= 1e-15 and abs(u.max() - 1) <= 1e-2
    ref = oracle.em_fit(data, 3, seed=42, means_init=oracle.KPP, absolute_tolerance=1e-10, relative_tolerance=0.0, maximum_steps=1000)
    assert abs(em.number_iterations - ref.iterations) <= 2   # a 1e-10 absolute tolerance on ll sits at rounding noise
    assert abs(em.log_likelihood - ref.log_likelihood) <= 1e-10


@pytest.mark.parametrize("kind", [oracle.FORGY, oracle.RANDOM_PARTITION, oracle.KPP])
def test_kmeans_two_gaussians(clustering, kind):
    """Tests/test_KMeans.cpp:8-91."""
    data, truth = oracle.testdata_two_gaussians()
    km = clustering.KMeans(2)
    km.set_absolute_tolerance(1e-8)
    km.set_maximum_steps(100)
    km.set_seed(63413131)
    km.set_centroids_initialiser(initialiser(clustering, kind))
    assert km.fit(data)
    labels = np.array(km.labels)
    assert labels.shape == (400,) and km.centroids.shape == (2, 3)
    total = 0.0
    for i in range(400):
        label, sq = km.assign_label(data[i])
        assert label == labels[i]
        total += sq
    assert abs(total - km.inertia) <= 1e-13
    centroids = km.centroids.T
    means = MEANS_TRUE.copy()
    if labels[0] != truth[0]:
        means = means[:, ::-1]
        truth = 1 - truth
    assert np.linalg.norm(means - centroids) <= 2e-2
    assert np.array_equal(labels, truth)
    ref = oracle.kmeans_fit(data, 2, seed=63413131, init=kind, maximum_steps=100)
    assert km.number_iterations == ref.iterations and km.inertia == pytest.approx(ref.inertia, rel=1e-12)
    km3 = clustering.KMeans(2)
    km3.set_seed(63413131)
    km3.set_centroids_initialiser(initialiser(clustering, kind))
    km3.set_number_initialisations(3)
    assert km3.fit(data)
    assert km3.inertia <= km.inertia + 1e-12
    km1 = clustering.KMeans(1)
    km1.set_centroids_initialiser(initialiser(clustering, kind))
    km1.fit(data)
    assert np.linalg.norm(data.mean(axis=0) - km1.centroids[0]) <= 1e-14
    assert set(km1.labels) == {0}


def test_kmeans_mouse(clustering):
    """cppyml/tests/test_clustering.py:76-95."""
    data = mouse_numpy()
    km = clustering.KMeans(3)
    km.set_seed(42)
    km.set_centroids_initialiser(clustering.KPP())
    km.set_number_initialisations(10)
    km.set_absolute_tolerance(1e-10)
    km.set_maximum_steps(1000)
    assert km.fit(data)
    assert km.inertia > 0
    assert set(km.labels) == {0, 1, 2}
    assert km.centroids.shape == (3, 2)
    for i in range(3):
        label, sq = km.assign_label(km.centroids[i])
        assert label == i and sq == 0
    ref = oracle.kmeans_fit(data, 3, seed=42, init=oracle.KPP, number_initialisations=10, absolute_tolerance=1e-10, maximum_steps=1000)
    assert km.inertia == pytest.approx(ref.inertia, rel=1e-12)
    assert np.array_equal(np.array(km.labels, dtype=np.uint32), ref.labels)


def test_em_seeded_forgy_on_a_larger_mixture(clustering):
    """Seed-for-seed with the oracle at a bench-like shape (D=8, K=16): same iterations, labels, parameters."""
    data, _, _ = synthetic_gmm(40000, 8, 16, seed=77, spread=8.0)
    em = clustering.EM(16)
    em.set_seed(2024)
    em.set_maximum_steps(400)
    converged = em.fit(data)
    ref = oracle.em_fit(data, 16, seed=2024, maximum_steps=400)
    assert converged == ref.converged and em.number_iterations == ref.iterations
    assert abs(em.log_likelihood - ref.log_likelihood) <= 1e-9 * abs(ref.log_likelihood)
    assert np.max(np.abs(em.means - ref.means)) <= 1e-9 * np.max(np.abs(ref.means))
    for k in range(16):
        assert np.max(np.abs(em.covariance(k) - ref.covariances[k])) <= 1e-9 * np.max(np.abs(ref.covariances[k]))
    assert np.max(np.abs(em.mixing_probabilities - ref.mixing_probabilities)) <= 1e-9
    assert np.max(np.abs(em.responsibilities - ref.responsibilities)) <= 1e-9
