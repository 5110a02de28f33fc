"""Pins the CPU oracle against the reference's own tests.

The reference has no stored golden vectors; its tests are property checks on data drawn inside each
test.  The oracle regenerates exactly the same data (same libstdc++ / numpy calls) and must pass
exactly the same assertions.  Runs without a GPU.
"""
import warnings

import numpy as np
import pytest

import oracle
from tests.datasets import mouse_numpy

MEANS_TRUE = np.array([[0.4, -1.2], [0.11, 2.2], [0.5, 1.6]])
SIGMAS_TRUE = np.array([[0.05, 0.2], [0.04, 0.1], [0.01, 0.2]])
P0 = 0.25


def check_two_gaussians_em(fit_fn, assign_fn, init, maximise_first):
    """Tests/test_EM.cpp:8-104 with `fit_fn(data, k, **kw)` as the implementation under test."""
    data, _ = oracle.testdata_two_gaussians()
    kw = dict(absolute_tolerance=1e-8, relative_tolerance=1e-8, maximum_steps=100, seed=63413131,
              maximise_first=maximise_first)
    if init is not None:
        kw["means_init"] = init
    em = fit_fn(data, 2, **kw)
    assert em.converged                                              # test_EM.cpp:48-49
    assert em.mixing_probabilities.shape == (2,)
    assert em.labels.shape == (400,)
    assert em.means.shape == (3, 2)
    assert em.responsibilities.shape == (400, 2)
    for i in range(400):                                              # test_EM.cpp:58-62
        u = assign_fn(em, data[i])
        assert np.linalg.norm(u - em.responsibilities[i]) <= 1e-15, i
    means = MEANS_TRUE.copy()
    covs = [np.diag(SIGMAS_TRUE[:, k] ** 2) for k in range(2)]
    mix = np.array([P0, 1 - P0])
    if (em.mixing_probabilities[0] < em.mixing_probabilities[1]) != (P0 < 1 - P0):   # test_EM.cpp:77-82
        mix = mix[::-1]
        means = means[:, ::-1]
        covs = covs[::-1]
    assert np.linalg.norm(mix - em.mixing_probabilities) <= 2e-2     # test_EM.cpp:83
    assert np.linalg.norm(means - em.means) <= 2e-2                  # test_EM.cpp:84
    for k in range(2):
        assert np.linalg.norm(covs[k] - em.covariances[k]) <= 1e-2   # test_EM.cpp:85-87
    kw1 = dict(maximise_first=maximise_first)
    if init is not None:
        kw1["means_init"] = init
    em1 = fit_fn(data, 1, **kw1)                                      # test_EM.cpp:89-103
    assert em1.log_likelihood <= em.log_likelihood
    assert np.linalg.norm(data.mean(axis=0) - em1.means[:, 0]) <= 1e-14
    for i in range(400):
        u = assign_fn(em1, data[i])
        assert np.linalg.norm(u - em1.responsibilities[i]) <= 1e-15
        assert em1.labels[i] == 0
    return em


@pytest.mark.parametrize("init,maximise_first", [(oracle.FORGY, False), (oracle.RANDOM_PARTITION, False),
                                                  (oracle.KPP, False), (None, True)])
def test_em_two_gaussians(init, maximise_first):
    check_two_gaussians_em(oracle.em_fit, oracle.em_assign_responsibilities, init, maximise_first)


def test_em_deterministic():
    """Tests/test_EM.cpp:126-144: N == K gives an exact fit."""
    data = np.array([[-1, 1, 0.5], [0, 0.5, 0.5]])
    em = oracle.em_fit(data, 2)
    assert em.converged
    assert em.log_likelihood == np.inf
    for i in range(2):
        assert em.labels[i] == i
        assert np.array_equal(em.means[:, i], data[i])


def test_em_argument_errors():
    """EM.cpp:96-101: N < K is std::invalid_argument."""
    with pytest.raises(ValueError):
        oracle.em_fit(np.zeros((2, 3)), 3)


def check_two_gaussians_kmeans(fit_fn, init, inertia_tol=1e-15):
    """Tests/test_KMeans.cpp:8-91."""
    data, truth = oracle.testdata_two_gaussians()
    kw = dict(absolute_tolerance=1e-8, maximum_steps=100, seed=63413131)
    if init is not None:
        kw["init"] = init
    km = fit_fn(data, 2, **kw)
    assert km.converged
    assert km.centroids.shape == (3, 2)
    assert km.labels.shape == (400,)
    inertia = 0.0
    for i in range(400):                                              # test_KMeans.cpp:57-62
        label, sq = oracle.kmeans_assign_label(km.centroids, data[i])
        assert label == km.labels[i]
        assert abs(np.sum((km.centroids[:, label] - data[i]) ** 2) - sq) <= 1e-15
        inertia += sq
    assert abs(inertia - km.inertia) <= inertia_tol                   # test_KMeans.cpp:63
    truth = truth.copy()
    centroids = MEANS_TRUE.copy()
    if truth[0] != km.labels[0]:                                      # test_KMeans.cpp:66-71
        truth = 1 - truth
        centroids = centroids[:, ::-1]
    assert np.linalg.norm(centroids - km.centroids) <= 2e-2
    assert np.array_equal(truth, km.labels)                           # test_KMeans.cpp:73
    km3 = fit_fn(data, 2, number_initialisations=3, **kw)             # test_KMeans.cpp:75-79
    assert km3.converged
    assert km3.inertia <= inertia
    kw1 = {} if init is None else {"init": init}
    km1 = fit_fn(data, 1, **kw1)                                      # test_KMeans.cpp:81-90
    assert np.linalg.norm(data.mean(axis=0) - km1.centroids[:, 0]) <= 1e-14
    return km


@pytest.mark.parametrize("init", [oracle.FORGY, oracle.RANDOM_PARTITION, oracle.KPP])
def test_kmeans_two_gaussians(init):
    check_two_gaussians_kmeans(oracle.kmeans_fit, init)


def test_kmeans_deterministic():
    """Tests/test_KMeans.cpp:108-127."""
    data = np.array([[-1, 1, 0.5], [0, 0.5, 0.5]])
    km = oracle.kmeans_fit(data, 2)
    assert km.converged
    assert km.inertia == 0
    for i in range(2):
        assert km.labels[i] == i
        assert np.array_equal(km.centroids[:, i], data[i])


def test_linear_algebra_identities():
    """Tests/test_LinearAlgebra.cpp:17-75: xAx, xxT and add_a_xxT against dense products."""
    rng = np.random.default_rng(5)
    for n in (4, 14, 15, 64, 1024):
        a = rng.standard_normal((n, n))
        a = a + a.T
        x = rng.standard_normal(n)
        expected = x @ a @ x
        assert abs(oracle.xAx_symmetric(a, x) - expected) <= 1e-14 * max(1.0, np.abs(a).sum() * np.abs(x).max() ** 2)
    for n in (3, 10, 11, 13, 14, 40):
        x = rng.standard_normal(n)
        assert np.allclose(oracle.xxT(x), np.outer(x, x), rtol=1e-15, atol=0)
        dest = rng.standard_normal((n, n))
        got = oracle.add_a_xxT(x, dest, -0.3)
        assert np.linalg.norm(got - (dest - 0.3 * np.outer(x, x))) <= 1e-15 * np.linalg.norm(got) * n


def test_em_mouse_matches_sklearn():
    """cppyml/tests/test_clustering.py:47-74: the log-likelihood equals sklearn's within 1e-10."""
    import sklearn.mixture
    data = mouse_numpy()
    em = oracle.em_fit(data, 3, seed=42, absolute_tolerance=1e-10, relative_tolerance=0, means_init=oracle.KPP,
                       maximum_steps=1000)
    assert em.converged
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        gmm = sklearn.mixture.GaussianMixture(3, tol=1e-10, max_iter=1000, random_state=999, n_init=1, reg_covar=1e-15)
        gmm.fit(data)
    assert abs(gmm.score(data) - em.log_likelihood) <= 1e-10
    u = oracle.em_assign_responsibilities(em, np.array([0.0, 0.0]))
    assert len(u) == 3
    assert abs(1 - u.sum()) <= 1e-15
    assert u.min() >= 0
    assert abs(1 - u.max()) <= 1e-9


def test_kmeans_mouse():
    """cppyml/tests/test_clustering.py:76-95."""
    data = mouse_numpy()
    km = oracle.kmeans_fit(data, 3, seed=42, absolute_tolerance=1e-10, init=oracle.KPP, maximum_steps=1000,
                           number_initialisations=10)
    assert km.converged
    assert km.inertia > 0
    assert km.labels.min() == 0 and km.labels.max() == 2
    assert km.centroids.shape == (2, 3)
    for i in range(3):
        label, sq = oracle.kmeans_assign_label(km.centroids, km.centroids[:, i])
        assert label == i and sq == 0
