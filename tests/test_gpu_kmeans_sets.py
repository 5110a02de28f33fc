"""K-means start sets (mlb_kms; SURVEY.md 8(f) row 3, KMeans.cpp:29-47): several starts of a multi-start fit advanced by ONE
pass of the assignment kernel over the points.  Per start the results must be bit for bit those of the start run alone on an
mlb_km (same kernels, same summation order), and through the host class the fit must be the reference's (the oracle's)."""
import numpy as np
import pytest

import oracle
from tests.datasets import synthetic_gmm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from ml_b200 import cabi as module
    assert module.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return module


@pytest.fixture(scope="module")
def ctx(cabi):
    c = cabi.Context(1)
    yield c
    c.close()


def _starts(data, k, n_sets, seed):
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray(data[rng.choice(len(data), size=k, replace=False)].T) for _ in range(n_sets)]


@pytest.mark.parametrize("n,d,k,n_sets", [(30011, 8, 20, 3), (20000, 16, 48, 4), (9000, 32, 100, 2), (5000, 2, 3, 4), (12000, 5, 7, 3), (7000, 64, 40, 2)])
def test_start_sets_equal_the_starts_run_alone(cabi, ctx, n, d, k, n_sets):
    data, _, _ = synthetic_gmm(n, d, max(2, k // 2), seed=n % 97, spread=6.0)
    dev = cabi.Data.upload(ctx, data)
    assert cabi.Kms.supported(dev, k, n_sets)
    starts = _starts(data, k, n_sets, seed=d)
    alone = []
    for s in range(n_sets):
        km = cabi.Km(dev, k)
        km.set_centroids(starts[s])
        trace = []
        for _ in range(4):
            inertia, changed = km.assign()
            trace.append((inertia, changed, km.update()))
        inertia, changed = km.assign()
        alone.append((trace, inertia, changed, km.get_labels().copy(), km.get_centroids()))
        km.close()
    sets = cabi.Kms(dev, k, n_sets)
    for s in range(n_sets):
        sets.set_centroids(s, starts[s])
    for it in range(4):
        inertia, changed = sets.assign()
        shift = sets.update()
        for s in range(n_sets):
            # labels (hence changed counts, statistics, centroids and their shift) are bit for bit the start's own; a chunk's
            # inertia partial is summed per warp, so it is exact only when both objects run the same number of warps per CTA
            # (they do whenever the sets' centroid images leave room for it)
            assert (changed[s], shift[s]) == alone[s][0][it][1:], (it, s)
            assert abs(inertia[s] - alone[s][0][it][0]) <= 1e-14 * inertia[s], (it, s)
    inertia, changed = sets.assign()
    for s in range(n_sets):
        assert abs(inertia[s] - alone[s][1]) <= 1e-14 * inertia[s] and changed[s] == alone[s][2], s
        assert np.array_equal(sets.get_labels(s), alone[s][3]), s
        assert np.array_equal(sets.get_centroids(s), alone[s][4]), s
    sets.close(); dev.close()


def test_frozen_start_keeps_its_state_while_the_others_advance(cabi, ctx):
    n, d, k = 25000, 8, 24
    data, _, _ = synthetic_gmm(n, d, 12, seed=5, spread=6.0)
    dev = cabi.Data.upload(ctx, data)
    starts = _starts(data, k, 3, seed=11)
    sets = cabi.Kms(dev, k, 3)
    for s in range(3):
        sets.set_centroids(s, starts[s])
    for _ in range(2):
        sets.assign()
        sets.update()
    sets.assign()
    frozen_labels, frozen_centroids = sets.get_labels(1).copy(), sets.get_centroids(1)
    active = 0b101
    for _ in range(3):
        sets.update(active)
        inertia, changed = sets.assign(active)
        assert inertia[1] == 0 and changed[1] == 0      # not reported for a start outside the mask
    assert np.array_equal(sets.get_labels(1), frozen_labels) and np.array_equal(sets.get_centroids(1), frozen_centroids)
    # the active starts went on exactly as they do alone
    for s in (0, 2):
        km = cabi.Km(dev, k)
        km.set_centroids(starts[s])
        for _ in range(5):
            km.assign()
            km.update()
        km.assign()
        assert np.array_equal(sets.get_labels(s), km.get_labels()), s
        assert np.array_equal(sets.get_centroids(s), km.get_centroids()), s
        km.close()
    sets.close(); dev.close()


def test_shapes_that_do_not_fit_are_reported(cabi, ctx):
    data, _, _ = synthetic_gmm(3000, 96, 4, seed=2)
    dev = cabi.Data.upload(ctx, data)
    assert not cabi.Kms.supported(dev, 8, 2)            # D > 64: the exact scan runs one start at a time
    with pytest.raises(cabi.MlbError):
        cabi.Kms(dev, 8, 2)
    dev.close()
    data, _, _ = synthetic_gmm(3000, 32, 4, seed=2)
    dev = cabi.Data.upload(ctx, data)
    assert cabi.Kms.supported(dev, 256, 2) and not cabi.Kms.supported(dev, 256, 4)   # four 64 KB images do not fit one CTA
    assert not cabi.Kms.supported(dev, 8, 5)
    dev.close()


@pytest.mark.parametrize("n_inits,max_steps", [(3, 100), (6, 100), (2, 100), (3, 2), (5, 3)])
def test_multi_start_fit_through_cppyml_matches_oracle(n_inits, max_steps):
    """KMeans::fit with several initialisations (KMeans.cpp:29-47; the 3-start K-means++ fit of Benchmarks/bm_KMeans.cpp) runs
    its starts in lockstep groups of four; the result is the reference's: best converged start by inertia, or the last start's
    state when none converged within maximum_steps."""
    from ml_b200 import import_cppyml
    cppyml = import_cppyml()
    data, _ = oracle.testdata_mouse(4000)
    km = cppyml.clustering.KMeans(3)
    km.set_seed(77)
    km.set_absolute_tolerance(1e-14)
    km.set_centroids_initialiser(cppyml.clustering.KPP())
    km.set_number_initialisations(n_inits)
    km.set_maximum_steps(max_steps)
    converged = km.fit(data)
    ref = oracle.kmeans_fit(data, 3, init=oracle.KPP, seed=77, absolute_tolerance=1e-14, number_initialisations=n_inits, maximum_steps=max_steps)
    assert converged == ref.converged
    assert np.array_equal(np.asarray(km.labels, dtype=np.uint32), ref.labels)
    assert abs(km.inertia - ref.inertia) <= 1e-12 * max(ref.inertia, 1e-300)
    assert np.max(np.abs(np.asarray(km.centroids) - ref.centroids.T)) <= 1e-12
