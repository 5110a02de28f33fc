"""SURVEY.md §8(f) rows 1 and 2 on the GPU: the initialisers' distance passes (KPP::init and ClosestCentroid::init,
ML/Clustering.cpp:39-89) and batched prediction (EM::assign_responsibilities, ML/EM.cpp:176-188; KMeans::assign_label,
ML/KMeans.cpp:153-165), each against the CPU oracle on the same inputs.

Bars: labels, drawn centroids and iteration counts identical; squared distances within 1e-13 relative (the same
fused-multiply-add chain, compared loosely only because the host compiler is free to vectorise the oracle's loop);
responsibilities within 1e-9 absolute, parameters within 1e-9 relative (as in test_gpu_cabi_parity.py).
"""
import numpy as np
import pytest

import oracle
from tests.datasets import mouse_numpy, synthetic_gmm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from ml_b200 import cabi as module
    assert module.device_count() >= 1
    return module


@pytest.fixture(scope="module")
def ctx(cabi):
    c = cabi.Context(1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def clustering():
    from ml_b200 import import_cppyml
    return import_cppyml().clustering


# ---------------------------------------------------------------- KPP distance pass (Clustering.cpp:42-51)

@pytest.mark.parametrize("n,d", [(1, 1), (127, 3), (128, 2), (129, 5), (1000, 13), (4097, 16), (30000, 32), (5000, 64), (777, 100)])
def test_kpp_distance_pass(cabi, ctx, n, d):
    rng = np.random.default_rng(n * 131 + d)
    data = np.ascontiguousarray(rng.normal(size=(n, d)) * 3.0 + 1.5)
    dev = cabi.Data.upload(ctx, data)
    picks = [int(i) for i in rng.integers(0, n, size=4)]
    nearest_ref = np.full(n, np.inf)
    for j, index in enumerate(picks):
        # the reference recomputes min_k |x - c_k|^2 from scratch (Clustering.cpp:45-50); the device folds in the newest
        nearest_ref = np.minimum(nearest_ref, np.array([oracle.kmeans_assign_label(data[index][:, None], x)[1] for x in data]) if n <= 1000
                                 else ((data - data[index]) ** 2).sum(axis=1))
        nearest = dev.kpp_update(data[index], first=(j == 0))
        assert nearest.shape == (n,)
        scale = np.maximum(nearest_ref, 1e-300)
        assert np.max(np.abs(nearest - nearest_ref) / scale) <= (1e-13 if n <= 1000 else 1e-12), (n, d, j)
        assert nearest[index] == 0.0
    assert dev.launch_count == len(picks)
    dev.close()


def test_kpp_update_needs_a_first_pass(cabi, ctx):
    data = np.ascontiguousarray(np.random.default_rng(1).normal(size=(50, 3)))
    dev = cabi.Data.upload(ctx, data)
    with pytest.raises(cabi.MlbError) as e:
        dev.kpp_update(data[0], first=False)
    assert e.value.code == cabi.MLB_EINVAL
    dev.close()


@pytest.mark.parametrize("n,d,k,seed", [(400, 2, 2, 63413131), (3000, 2, 3, 42), (20000, 8, 16, 7), (5000, 20, 6, 11), (2500, 33, 4, 5)])
def test_kpp_seeded_fits_match_oracle(clustering, n, d, k, seed):
    """A seeded fit whose means come from the built-in KPP (device distance passes, host draws) is the oracle's fit:
    the same centroids were drawn, in the same order, from the same PRNG stream."""
    data = mouse_numpy(n) if d == 2 and n == 3000 else (oracle.testdata_two_gaussians()[0] if n == 400 else synthetic_gmm(n, d, k, seed=seed, spread=6.0)[0])
    km = clustering.KMeans(k)
    km.set_seed(seed)
    km.set_centroids_initialiser(clustering.KPP())
    km.set_maximum_steps(300)
    km.fit(data)
    ref = oracle.kmeans_fit(data, k, seed=seed, init=oracle.KPP, maximum_steps=300)
    assert km.number_iterations == ref.iterations and km.converged == ref.converged
    assert np.array_equal(np.asarray(km.labels, dtype=np.uint32), ref.labels)
    assert abs(km.inertia - ref.inertia) <= 1e-9 * ref.inertia
    assert np.max(np.abs(km.centroids.T - ref.centroids)) <= 1e-9 * max(1.0, np.max(np.abs(ref.centroids)))

    em = clustering.EM(k)
    em.set_seed(seed)
    em.set_means_initialiser(clustering.KPP())
    em.set_maximum_steps(60)
    em.fit(data)
    eref = oracle.em_fit(data, k, seed=seed, means_init=oracle.KPP, maximum_steps=60)
    assert em.number_iterations == eref.iterations
    assert abs(em.log_likelihood - eref.log_likelihood) <= 1e-9 * abs(eref.log_likelihood)
    assert np.max(np.abs(em.means - eref.means)) <= 1e-9 * max(1.0, np.max(np.abs(eref.means)))


def test_kpp_multi_start_kmeans_matches_oracle(clustering):
    """KMeans.cpp:29-47 with KPP: every initialisation restarts the resident `nearest` vector (first pass again)."""
    data, _, _ = synthetic_gmm(6000, 5, 7, seed=21, spread=5.0)
    km = clustering.KMeans(7)
    km.set_seed(5)
    km.set_centroids_initialiser(clustering.KPP())
    km.set_number_initialisations(4)
    km.fit(data)
    ref = oracle.kmeans_fit(data, 7, seed=5, init=oracle.KPP, number_initialisations=4)
    assert km.converged == ref.converged
    assert np.array_equal(np.asarray(km.labels, dtype=np.uint32), ref.labels)
    assert abs(km.inertia - ref.inertia) <= 1e-9 * ref.inertia


# ---------------------------------------------------------------- ClosestCentroid start (Clustering.cpp:72-89)

@pytest.mark.parametrize("n,d,k", [(5000, 2, 3), (20000, 8, 16), (6000, 16, 32), (3000, 20, 5), (4000, 12, 40)])
def test_mstep_from_labels_equals_one_hot_responsibilities(cabi, ctx, n, d, k):
    """The M-step from hard labels is bit for bit the M-step of the one-hot matrix ClosestCentroid::init writes."""
    data, labels, _ = synthetic_gmm(n, d, k, seed=n + d, spread=6.0)
    dev = cabi.Data.upload(ctx, data)
    onehot = np.zeros((n, k))
    onehot[np.arange(n), labels] = 1.0
    a = cabi.Em(dev, k)
    a.mstep_from_responsibilities(onehot)
    b = cabi.Em(dev, k)
    b.mstep_from_labels(labels)
    for x, y in zip(a.get_params(), b.get_params()):
        assert np.array_equal(x, y)
    with pytest.raises(cabi.MlbError):
        bad = labels.copy()
        bad[n // 2] = k
        b.mstep_from_labels(bad)
    a.close(), b.close(), dev.close()


@pytest.mark.parametrize("n,d,k,centroids_kind,spread", [(400, 2, 2, oracle.FORGY, 0), (10000, 8, 16, oracle.FORGY, 7.0), (4000, 6, 5, oracle.KPP, 7.0),
                                                         (3000, 4, 5, oracle.RANDOM_PARTITION, 7.0), (3000, 20, 5, oracle.RANDOM_PARTITION, 2.0),
                                                         (3000, 20, 5, oracle.FORGY, 7.0), (16000, 8, 36, oracle.FORGY, 7.0)])
def test_closest_centroid_start_matches_oracle(clustering, n, d, k, centroids_kind, spread):
    """set_maximise_first(True): initial centroids on the host PRNG, nearest-centroid pass and one-hot M-step on the device.
    (Shapes where every initial cluster holds several times D points: a cluster of fewer than D points starts from a
    singular covariance + 1e-15 I, and the reference's own arithmetic is rounding noise from there on.)"""
    data = oracle.testdata_two_gaussians()[0] if n == 400 else synthetic_gmm(n, d, k, seed=3 * n + d, spread=spread)[0]
    kinds = {oracle.FORGY: clustering.Forgy, oracle.RANDOM_PARTITION: clustering.RandomPartition, oracle.KPP: clustering.KPP}
    em = clustering.EM(k)
    em.set_seed(977)
    em.set_maximise_first(True)
    em.set_responsibilities_initialiser(clustering.ClosestCentroid(kinds[centroids_kind]()))
    em.set_maximum_steps(50)
    em.fit(data)
    ref = oracle.em_fit(data, k, seed=977, maximise_first=True, resp_init_centroids=centroids_kind, maximum_steps=50)
    assert np.isfinite(ref.log_likelihood)   # parity is defined where the reference is finite (it has no log-sum-exp)
    assert em.number_iterations == ref.iterations
    assert abs(em.log_likelihood - ref.log_likelihood) <= 1e-9 * abs(ref.log_likelihood)
    assert np.max(np.abs(em.means - ref.means)) <= 1e-9 * max(1.0, np.max(np.abs(ref.means)))
    assert np.max(np.abs(em.mixing_probabilities - ref.mixing_probabilities)) <= 1e-9
    for c in range(k):
        assert np.max(np.abs(em.covariance(c) - ref.covariances[c])) <= 1e-9 * max(1.0, np.max(np.abs(ref.covariances[c])))


# ---------------------------------------------------------------- batched EM::assign_responsibilities (EM.cpp:176-188)

@pytest.mark.parametrize("n,d,k", [(3000, 2, 3), (20000, 8, 16), (8000, 16, 32), (4000, 20, 5), (5000, 12, 40), (3000, 64, 8)])
def test_em_predict_matches_oracle(cabi, ctx, n, d, k):
    """Fused (D <= 16, K <= 32) and split shapes: the device responsibilities of new points under the post-fit
    parameters against the oracle's assign_responsibilities with the oracle's post-fit parameters."""
    data, _, true_means = synthetic_gmm(n, d, k, seed=n + 7 * d, spread=8.0)
    # K data points at fixed indices; in many dimensions the generating means (see test_gpu_cabi_parity.py)
    init = np.ascontiguousarray(true_means.T if d >= 48 else data[:: n // k][:k].T)
    steps = 5
    ref = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps,
                        absolute_tolerance=0.0, relative_tolerance=0.0, want_responsibilities=False)
    assert ref.iterations == steps
    dev = cabi.Data.upload(ctx, data)
    em = cabi.Em(dev, k)
    with pytest.raises(cabi.MlbError) as e:
        em.predict(data[:4])
    assert e.value.code == cabi.MLB_ESTATE
    em.set_params(init, np.repeat(em.sample_covariance()[None], k, axis=0), np.full(k, 1.0 / k))
    em.run_steps(steps)
    rng = np.random.default_rng(5)
    queries = np.ascontiguousarray(np.concatenate([data[rng.integers(0, n, size=150)], data[:50] + 0.3 * rng.normal(size=(50, d))]))
    resp, labels = em.predict(queries)
    assert resp.shape == (200, k) and labels.shape == (200,)
    expect = np.array([oracle.em_assign_responsibilities(ref, q) for q in queries])
    assert np.max(np.abs(resp - expect)) <= 1e-9
    assert np.max(np.abs(resp.sum(axis=1) - 1.0)) <= 1e-12
    clear = np.sort(expect, axis=1)[:, -1] - np.sort(expect, axis=1)[:, -2] > 1e-6 if k > 1 else np.ones(200, dtype=bool)
    assert np.array_equal(labels[clear], np.argmax(expect, axis=1)[clear].astype(np.uint32))
    # ragged sizes and either output alone
    for m in (1, 63, 65, 129):
        r_only, none = em.predict(queries[:m], want_labels=False)
        assert none is None and np.array_equal(r_only, resp[:m])
        none, l_only = em.predict(queries[:m], want_responsibilities=False)
        assert none is None and np.array_equal(l_only, labels[:m])
    em.close(), dev.close()


def test_em_predict_is_staged_consistently(cabi, ctx):
    """More points than one staging batch (2^20): every batch gives what a direct call on its rows gives."""
    data, _, _ = synthetic_gmm(4000, 2, 3, seed=1, spread=5.0)
    dev = cabi.Data.upload(ctx, data)
    em = cabi.Em(dev, 3)
    em.set_params(data[:3].T, np.repeat(em.sample_covariance()[None], 3, axis=0), np.full(3, 1.0 / 3))
    em.run_steps(5)
    m = (1 << 20) + 12345
    queries = np.ascontiguousarray(np.random.default_rng(9).normal(size=(m, 2)) * 4.0)
    resp, labels = em.predict(queries)
    for lo in (0, (1 << 20) - 70, (1 << 20), m - 100):
        r, l = em.predict(queries[lo:lo + 100])
        assert np.array_equal(r, resp[lo:lo + 100]) and np.array_equal(l, labels[lo:lo + 100])
    em.close(), dev.close()


def test_cppyml_assign_responsibilities_batch(clustering):
    data, _, _ = synthetic_gmm(5000, 4, 3, seed=12, spread=5.0)
    em = clustering.EM(3)
    em.set_seed(3)
    em.fit(data)
    batch = em.assign_responsibilities_batch(np.ascontiguousarray(data[:300]))
    assert batch.shape == (300, 3)
    single = np.array([em.assign_responsibilities(x) for x in data[:300]])
    assert np.max(np.abs(batch - single)) <= 1e-9
    with pytest.raises(ValueError):
        em.assign_responsibilities_batch(np.zeros((5, 7)))
    with pytest.raises(TypeError):
        em.assign_responsibilities_batch(np.zeros((5, 4), dtype=np.float32))


# ---------------------------------------------------------------- batched KMeans::assign_label (KMeans.cpp:153-165)

@pytest.mark.parametrize("n,d,k", [(2000, 2, 3), (20000, 8, 16), (9000, 32, 256), (3000, 13, 33), (2000, 64, 20), (500, 1, 2)])
def test_km_predict_matches_oracle(cabi, ctx, n, d, k):
    data, _, _ = synthetic_gmm(n, d, max(2, k // 8), seed=n + d, spread=6.0)
    dev = cabi.Data.upload(ctx, data)
    km = cabi.Km(dev, k)
    with pytest.raises(cabi.MlbError) as e:
        km.predict(data[:4])
    assert e.value.code == cabi.MLB_ESTATE
    km.set_centroids(data[:k].T)
    for _ in range(3):
        km.assign()
        km.update()
    inertia, _ = km.assign()
    fit_labels = km.get_labels()
    centroids = km.get_centroids()
    rng = np.random.default_rng(3)
    queries = np.ascontiguousarray(np.concatenate([data[:137], data[rng.integers(0, n, size=100)] + rng.normal(size=(100, d))]))
    labels, dist = km.predict(queries)
    expect = [oracle.kmeans_assign_label(centroids, q) for q in queries]
    assert np.array_equal(labels, np.array([e[0] for e in expect], dtype=np.uint32))
    expect_d = np.array([e[1] for e in expect])
    assert np.max(np.abs(dist - expect_d) / np.maximum(expect_d, 1e-300)) <= 1e-13
    # prediction on the training points is the fit's own assignment, and leaves the fit untouched
    all_labels, all_dist = km.predict(data)
    assert np.array_equal(all_labels, fit_labels)
    assert abs(all_dist.sum() - inertia) <= 1e-11 * inertia
    assert np.array_equal(km.get_labels(), fit_labels) and np.array_equal(km.get_centroids(), centroids)
    km.close(), dev.close()


def test_cppyml_assign_labels(clustering):
    data, _, _ = synthetic_gmm(4000, 3, 5, seed=2, spread=6.0)
    km = clustering.KMeans(5)
    km.set_seed(8)
    km.fit(data)
    labels, dist = km.assign_labels(data)
    assert labels.dtype == np.uint32 and np.array_equal(labels, np.asarray(km.labels, dtype=np.uint32))
    assert abs(dist.sum() - km.inertia) <= 1e-11 * km.inertia
    for i in range(0, 4000, 400):
        label, sq = km.assign_label(data[i])
        assert label == labels[i] and abs(sq - dist[i]) <= 1e-13 * max(sq, 1e-300)
    with pytest.raises(ValueError):
        km.assign_labels(np.zeros((5, 4)))
