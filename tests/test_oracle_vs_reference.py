"""The CPU restatement (oracle/mlpp_oracle.cpp) against the reference's OWN code.

oracle/_ref/libmlpp_ref.so is built by oracle/Makefile from the reference's four clustering translation units where they
lie under /root/reference (ML/EM.cpp, ML/KMeans.cpp, ML/Clustering.cpp, ML/LinearAlgebra.cpp), with the first-party
Eigen stand-in of oracle/eigen_standin/ in place of the Eigen 3 headers the image lacks.  Control flow, loop structure,
PRNG use and scalar arithmetic are therefore the reference's; only the evaluation order INSIDE Eigen's dense products,
reductions and LLT is the stand-in's (sequential).  The restatement makes the same choice, so the two agree operation
for operation: compiled with -ffp-contract=off they are bit-identical (test_bitwise_without_fma_contraction); with the
reference's release flags the compiler fuses a*b+c differently in the two sources, which moves results by an ulp, so
the default builds are compared at 1e-13 relative with identical iteration counts, convergence flags and labels.

Skipped when the prebuilt library is absent (it cannot be rebuilt without /root/reference)."""
import numpy as np
import pytest

import oracle
from tests.datasets import mouse_numpy, synthetic_gmm

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libmlpp_ref.so not built (needs /root/reference)")


def close(x, y, exact):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if exact:
        return np.array_equal(x, y, equal_nan=True)
    return bool(np.all(np.abs(x - y) <= 1e-13 * max(1e-300, float(np.max(np.abs(y))))))


def same_em(a, b, exact=False):
    assert a.converged == b.converged
    assert a.iterations == b.iterations
    assert close(a.log_likelihood, b.log_likelihood, exact) or (np.isinf(a.log_likelihood) and a.log_likelihood == b.log_likelihood)
    assert close(a.means, b.means, exact)
    assert close(a.covariances, b.covariances, exact)
    assert close(a.mixing_probabilities, b.mixing_probabilities, exact)
    assert close(a.responsibilities, b.responsibilities, exact)
    if a.converged:
        assert np.array_equal(a.labels, b.labels)


def same_kmeans(a, b, exact=False):
    assert a.converged == b.converged
    assert a.iterations == b.iterations
    assert close(a.inertia, b.inertia, exact)
    assert close(a.centroids, b.centroids, exact)
    assert np.array_equal(a.labels, b.labels)


@pytest.mark.parametrize("init", [oracle.FORGY, oracle.RANDOM_PARTITION, oracle.KPP])
@pytest.mark.parametrize("maximise_first", [False, True])
def test_em_on_the_reference_test_data(init, maximise_first):
    """Tests/test_EM.cpp:36-87: the 400-point two-Gaussian data, every initialiser, both start modes."""
    data, _ = oracle.testdata_two_gaussians()
    kw = dict(seed=42, means_init=init, resp_init_centroids=init, maximise_first=maximise_first, maximum_steps=100)
    a = oracle.em_fit(data, 2, **kw)
    b = oracle.em_fit(data, 2, impl="reference", queries=data[:16], **kw)
    same_em(a, b)
    # EM::assign_responsibilities of the reference (EM.cpp:176-188) against the oracle's, from the post-fit parameters
    for i in range(16):
        assert close(oracle.em_assign_responsibilities(a, data[i]), b.query_responsibilities[i], False)


def test_em_mouse_benchmark_configuration():
    """BASELINE config 1 (Benchmarks/bm_EM.cpp:9-48): mouse data, N = 10k, K = 3, KPP, tolerances 1e-14."""
    data, _ = oracle.testdata_mouse(10000)
    kw = dict(seed=42, means_init=oracle.KPP, absolute_tolerance=1e-14, relative_tolerance=1e-14)
    same_em(oracle.em_fit(data, 3, **kw), oracle.em_fit(data, 3, impl="reference", **kw))


@pytest.mark.parametrize("n,d,k,seed", [(3000, 2, 3, 1), (2000, 8, 5, 2), (1500, 13, 4, 3), (1500, 14, 4, 4), (1500, 15, 3, 5), (1200, 20, 4, 6),
                                        (900, 33, 3, 7)])
def test_em_on_synthetic_mixtures(n, d, k, seed):
    """Dimensions on both sides of the reference's hand-written / Eigen switches (LinearAlgebra.cpp:17, 59: 15 and 14)
    and of Eigen's blocked LLT (n >= 32 in the real library; unblocked in both stand-ins)."""
    data, _, _ = synthetic_gmm(n, d, k, seed=seed, spread=6.0)
    kw = dict(seed=7, means_init=oracle.FORGY, maximum_steps=40)
    same_em(oracle.em_fit(data, k, **kw), oracle.em_fit(data, k, impl="reference", **kw))


def test_em_explicit_means_plugin():
    """A user-supplied CentroidsInitialiser (Clustering.hpp:58-72) is how the parity tests fix the initial means."""
    data, _, _ = synthetic_gmm(2500, 6, 4, seed=11, spread=5.0)
    init = np.ascontiguousarray(data[::600][:4].T)
    kw = dict(means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=60)
    same_em(oracle.em_fit(data, 4, **kw), oracle.em_fit(data, 4, impl="reference", **kw))


def test_em_exact_fit_and_argument_errors():
    data = np.random.default_rng(0).normal(size=(5, 3))
    a, b = oracle.em_fit(data, 5), oracle.em_fit(data, 5, impl="reference")
    assert a.converged and b.converged and np.isinf(a.log_likelihood) and np.isinf(b.log_likelihood)
    assert np.array_equal(a.means, b.means) and np.array_equal(a.labels, b.labels) and np.array_equal(a.responsibilities, b.responsibilities)
    for impl in ("oracle", "reference"):
        with pytest.raises(ValueError):
            oracle.em_fit(data, 6, impl=impl)          # fewer points than components (EM.cpp:99-101)
        with pytest.raises(ValueError):
            oracle.em_fit(data, 0, impl=impl)          # no components (EM.cpp:34-36)


@pytest.mark.parametrize("init", [oracle.FORGY, oracle.RANDOM_PARTITION, oracle.KPP])
@pytest.mark.parametrize("inits", [1, 3])
def test_kmeans_on_the_reference_test_data(init, inits):
    """Tests/test_KMeans.cpp:40-73 plus the multi-start selection of KMeans.cpp:29-47."""
    data, _ = oracle.testdata_two_gaussians()
    kw = dict(seed=42, init=init, number_initialisations=inits, maximum_steps=100)
    a = oracle.kmeans_fit(data, 2, **kw)
    b = oracle.kmeans_fit(data, 2, impl="reference", queries=data[:32], **kw)
    if inits > 1:
        b.iterations = a.iterations   # not observable from outside for a multi-start fit
    same_kmeans(a, b)
    for i in range(32):
        label, sq = oracle.kmeans_assign_label(a.centroids, data[i])
        assert label == b.query_labels[i] and close(sq, b.query_distances[i], False)


@pytest.mark.parametrize("inits,max_steps", [(3, 2), (5, 3), (4, 100), (6, 100)])
def test_kmeans_multi_start_with_and_without_convergence(inits, max_steps):
    """KMeans.cpp:29-47 when some or all starts run out of steps: the best CONVERGED start wins; when none converged the
    object is left in the last start's state (labels of its last assignment, centroids of its last update).  The device
    path's lockstep multi-start (ml_b200/host/src/KMeans.cpp, fit_lockstep) is tested against the port on exactly these
    cases (tests/test_gpu_kmeans_sets.py), so the port is pinned to the reference's own code on them here."""
    data, _ = oracle.testdata_mouse(4000)
    kw = dict(seed=77, init=oracle.KPP, absolute_tolerance=1e-14, number_initialisations=inits, maximum_steps=max_steps)
    a = oracle.kmeans_fit(data, 3, **kw)
    b = oracle.kmeans_fit(data, 3, impl="reference", **kw)
    b.iterations = a.iterations   # not observable from outside for a multi-start fit
    same_kmeans(a, b)


@pytest.mark.parametrize("n,d,k,seed", [(4000, 2, 3, 1), (3000, 8, 16, 2), (2000, 32, 40, 3), (1000, 1, 4, 4)])
def test_kmeans_on_synthetic_mixtures(n, d, k, seed):
    data, _, _ = synthetic_gmm(n, d, min(k, 12), seed=seed, spread=5.0)
    kw = dict(seed=3, init=oracle.KPP, maximum_steps=200)
    same_kmeans(oracle.kmeans_fit(data, k, **kw), oracle.kmeans_fit(data, k, impl="reference", **kw))


def test_kmeans_mouse():
    data = mouse_numpy()
    kw = dict(seed=1, init=oracle.FORGY, absolute_tolerance=1e-14)
    same_kmeans(oracle.kmeans_fit(data, 3, **kw), oracle.kmeans_fit(data, 3, impl="reference", **kw))


@pytest.mark.parametrize("kind", [oracle.FORGY, oracle.RANDOM_PARTITION, oracle.KPP])
def test_initialisers_draw_the_same_prng_stream(kind):
    data, _, _ = synthetic_gmm(700, 5, 6, seed=21)
    for seed in (None, 0, 12345):
        assert close(oracle.centroids_init(kind, data, 6, seed=seed), oracle.centroids_init(kind, data, 6, seed=seed, impl="reference"), False)


@pytest.mark.parametrize("dim", [1, 2, 5, 10, 11, 13, 14, 15, 24, 40])
def test_linear_algebra_helpers(dim):
    """LinearAlgebra.cpp:8-73 on both sides of its size switches (15, 11, 14)."""
    rng = np.random.default_rng(dim)
    a = rng.normal(size=(dim, dim))
    a = a + a.T
    x = rng.normal(size=dim)
    assert abs(oracle.xAx_symmetric(a, x) - oracle.xAx_symmetric(a, x, impl="reference")) <= 1e-13 * np.abs(a).sum() * np.abs(x).max() ** 2
    assert close(oracle.xxT(x), oracle.xxT(x, impl="reference"), False)
    assert close(oracle.add_a_xxT(x, a, 0.37), oracle.add_a_xxT(x, a, 0.37, impl="reference"), False)


def test_bitwise_without_fma_contraction(tmp_path):
    """Both sources compiled with -ffp-contract=off: every arithmetic operation is then the one written in the source,
    and the restatement must reproduce the reference's own code BIT FOR BIT (EM and K-means, several shapes)."""
    import ctypes
    import os
    import subprocess
    reference = "/root/reference/ML"
    if not os.path.exists(os.path.join(reference, "EM.cpp")):
        pytest.skip("the reference checkout is not mounted here")
    here = os.path.dirname(os.path.abspath(oracle.__file__))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    flags = ["-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-w"]
    o_path, r_path = str(tmp_path / "oracle_nofma.so"), str(tmp_path / "ref_nofma.so")
    subprocess.check_call([cxx, *flags, "-o", o_path, os.path.join(here, "mlpp_oracle.cpp")])
    subprocess.check_call([cxx, *flags, "-I", os.path.join(here, "eigen_standin"), "-I", reference, "-o", r_path,
                           *[os.path.join(reference, f) for f in ("EM.cpp", "KMeans.cpp", "Clustering.cpp", "LinearAlgebra.cpp")],
                           os.path.join(here, "ref_shim.cpp")])
    saved = (oracle._lib, oracle._ref, oracle._LIB_PATH, oracle._REF_LIB_PATH, oracle.build)
    try:
        oracle._lib = oracle._ref = None
        oracle._LIB_PATH, oracle._REF_LIB_PATH = o_path, r_path
        oracle.build = lambda force=False: o_path
        mouse, _ = oracle.testdata_mouse(10000)
        kw = dict(seed=42, means_init=oracle.KPP, absolute_tolerance=1e-14, relative_tolerance=1e-14)
        same_em(oracle.em_fit(mouse, 3, **kw), oracle.em_fit(mouse, 3, impl="reference", **kw), exact=True)
        for n, d, k, seed in [(2000, 8, 5, 2), (1500, 14, 4, 4), (1500, 15, 3, 5), (1200, 20, 4, 6)]:
            data, _, _ = synthetic_gmm(n, d, k, seed=seed, spread=6.0)
            kw = dict(seed=7, means_init=oracle.FORGY, maximum_steps=40)
            same_em(oracle.em_fit(data, k, **kw), oracle.em_fit(data, k, impl="reference", **kw), exact=True)
            kw = dict(seed=3, init=oracle.KPP, maximum_steps=200)
            same_kmeans(oracle.kmeans_fit(data, k + 3, **kw), oracle.kmeans_fit(data, k + 3, impl="reference", **kw), exact=True)
    finally:
        oracle._lib, oracle._ref, oracle._LIB_PATH, oracle._REF_LIB_PATH, oracle.build = saved
