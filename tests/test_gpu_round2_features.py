"""Round-2 additions on the device: standardise_features (cppyml/cppyml/utils.py:8-28), lazy labels, ExplicitCentroids,
two models fitted from two host threads at once (the GIL is released inside fit), a K-means object created on the same
resident data as an EM object that already replays its step from a CUDA graph."""
import threading

import numpy as np
import pytest

import oracle
from tests.datasets import synthetic_gmm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cppyml():
    from ml_b200 import cabi, import_cppyml
    assert cabi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return import_cppyml()


@pytest.mark.parametrize("n,d,offset", [(1, 3, 0.0), (2, 1, 0.0), (1000, 5, 0.0), (40001, 16, 1e3), (300000, 8, -50.0), (5000, 300, 2.0)])
def test_standardise_features_matches_numpy(cppyml, n, d, offset):
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((n, d)) * rng.uniform(0.1, 30.0, size=d) + offset
    want = x.copy()
    want -= np.mean(want, axis=0)
    if n > 1:
        want /= np.std(want, axis=0, ddof=0)
    got = cppyml.utils.standardise_features(x)
    assert got.shape == x.shape and got is not x
    # 1e-15 relative to the scale of the standardised values (unit variance), plus the conditioning of x - mean itself
    tol = 1e-15 * (1.0 + abs(offset) / 0.1)
    assert np.max(np.abs(got - want)) <= tol * max(1.0, np.max(np.abs(want)))


def test_standardise_features_contract(cppyml):
    with pytest.raises(ValueError):
        cppyml.utils.standardise_features(np.zeros(5))
    empty = np.zeros((0, 4))
    assert cppyml.utils.standardise_features(empty).shape == (0, 4)
    x = np.arange(12.0).reshape(4, 3)
    keep = x.copy()
    cppyml.utils.standardise_features(x)
    assert np.array_equal(x, keep), "the input must not be modified"


def test_explicit_centroids_and_lazy_labels(cppyml):
    data, _, _ = synthetic_gmm(20000, 6, 5, seed=3, spread=8.0)
    init = np.ascontiguousarray(data[:5])
    km = cppyml.clustering.KMeans(5)
    km.set_centroids_initialiser(cppyml.clustering.ExplicitCentroids(init))
    km.fit(data)
    ref = oracle.kmeans_fit(data, 5, init=oracle.EXPLICIT, explicit_means=init.T)
    assert km.number_iterations == ref.iterations
    labels = km.labels_array
    assert labels.dtype == np.uint32 and np.array_equal(labels, ref.labels)
    assert km.labels == list(ref.labels)

    em = cppyml.clustering.EM(5)
    em.set_means_initialiser(cppyml.clustering.ExplicitCentroids(init))
    converged = em.fit(data)
    eref = oracle.em_fit(data, 5, means_init=oracle.EXPLICIT, explicit_means=init.T)
    assert converged == eref.converged and em.number_iterations == eref.iterations and em.converged == eref.converged
    assert np.array_equal(em.labels_array, eref.labels)
    resp = em.responsibilities
    assert not resp.flags.writeable and resp.shape == (20000, 5)
    assert np.max(np.abs(resp - eref.responsibilities)) <= 1e-9
    with pytest.raises(ValueError):
        bad = cppyml.clustering.EM(4)
        bad.set_means_initialiser(cppyml.clustering.ExplicitCentroids(init))
        bad.fit(data)


def test_two_threads_fit_two_models(cppyml):
    """fit() releases the GIL and every model shares the process-wide context (one stream, one set of bounce
    buffers, graph capture on that stream).  The context serialises its entry points; both fits must match the oracle."""
    data_a, _, _ = synthetic_gmm(60000, 8, 6, seed=11, spread=8.0)
    data_b, _, _ = synthetic_gmm(50000, 8, 4, seed=12, spread=8.0)
    out = {}

    def run(name, data, k, seed):
        for rep in range(3):
            em = cppyml.clustering.EM(k)
            em.set_seed(seed)
            em.set_maximum_steps(60)
            em.fit(data)
            km = cppyml.clustering.KMeans(k)
            km.set_seed(seed)
            km.fit(data)
            out[name] = (em.number_iterations, em.log_likelihood, np.array(em.means), km.inertia, km.number_iterations)

    ta = threading.Thread(target=run, args=("a", data_a, 6, 5))
    tb = threading.Thread(target=run, args=("b", data_b, 4, 7))
    ta.start(); tb.start(); ta.join(); tb.join()
    for name, data, k, seed in (("a", data_a, 6, 5), ("b", data_b, 4, 7)):
        ref = oracle.em_fit(data, k, seed=seed, maximum_steps=60)
        kref = oracle.kmeans_fit(data, k, seed=seed)
        iters, ll, means, inertia, kiters = out[name]
        assert iters == ref.iterations
        assert abs(ll - ref.log_likelihood) <= 1e-9 * abs(ref.log_likelihood)
        assert np.max(np.abs(means - ref.means)) <= 1e-9 * np.max(np.abs(ref.means))
        assert kiters == kref.iterations and abs(inertia - kref.inertia) <= 1e-9 * kref.inertia


def test_kmeans_on_shared_data_does_not_disturb_a_graphed_em():
    """the reduction scratch used to live in the shared data object, and a K-means object with a longer
    statistics vector reallocated it under the EM's captured step graph.  N is large enough for the two-level reduction."""
    from ml_b200 import cabi
    ctx = cabi.Context(1)
    data = cabi.Data.generate_gmm(ctx, 1_400_000, 8, 6, seed=77)   # 10938 chunks > 512: two-level reduction
    init = np.ascontiguousarray(data.download(0, 6).T)

    def run(interleave):
        em = cabi.Em(data, 6)
        cov = em.sample_covariance()
        em.set_params(init, np.repeat(cov[None], 6, axis=0), np.full(6, 1.0 / 6))
        lls = [em.step() for _ in range(4)]          # the third step captures the graph
        if interleave:
            km = cabi.Km(data, 200)                  # SV = 200 * 9 + 8 > the EM's: the old shared scratch would grow here
            km.set_centroids(np.ascontiguousarray(data.download(0, 200).T))
            km.assign()
            km.update()
        lls += [em.step() for _ in range(4)]
        params = em.get_params()
        em.close()
        if interleave:
            km.close()
        return lls, params

    plain, interleaved = run(False), run(True)
    assert plain[0] == interleaved[0]
    for a, b in zip(plain[1], interleaved[1]):
        assert np.array_equal(a, b)
    data.close()
    ctx.close()


def test_release_device_keeps_the_results(cppyml):
    data, _, _ = synthetic_gmm(8000, 4, 3, seed=9, spread=8.0)
    em = cppyml.clustering.EM(3)
    em.set_seed(3)
    em.fit(data)
    ref = oracle.em_fit(data, 3, seed=3)
    em.release_device()
    assert np.max(np.abs(em.responsibilities - ref.responsibilities)) <= 1e-9
    assert np.array_equal(em.labels_array, ref.labels)
    with pytest.raises(Exception):
        em.assign_responsibilities_batch(data[:10])
    km = cppyml.clustering.KMeans(3)
    km.set_seed(3)
    km.fit(data)
    kref = oracle.kmeans_fit(data, 3, seed=3)
    km.release_device()
    assert np.array_equal(km.labels_array, kref.labels)
