"""Golden fixtures (tests/golden/*.npz): outputs of the reference's own code (oracle/_ref) generated in the builder
container by tests/golden/make_golden.py.  CPU: the oracle port reproduces them.  GPU: the CUDA path, called through
the C-ABI, reproduces them to the north star's bar (identical iteration count, convergence flag and labels; parameters
and log-likelihood within 1e-9 relative)."""
import glob
import os

import numpy as np
import pytest

import oracle

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EM_FILES = sorted(glob.glob(os.path.join(HERE, "em_*.npz")))
KM_FILES = sorted(glob.glob(os.path.join(HERE, "km_*.npz")))
RTOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_fixtures_are_present():
    assert len(EM_FILES) >= 6 and len(KM_FILES) >= 4


@pytest.mark.parametrize("path", EM_FILES, ids=os.path.basename)
def test_oracle_reproduces_golden_em(path):
    g = np.load(path)
    k = g["initial_means"].shape[1]
    fit = oracle.em_fit(g["data"], k, means_init=oracle.EXPLICIT, explicit_means=g["initial_means"], maximum_steps=int(g["maximum_steps"]))
    assert fit.iterations == int(g["iterations"]) and fit.converged == bool(g["converged"])
    assert abs(fit.log_likelihood - float(g["log_likelihood"])) <= 1e-13 * abs(float(g["log_likelihood"]))
    assert rel(fit.means, g["means"]) <= 1e-13 and rel(fit.covariances, g["covariances"]) <= 1e-13
    assert rel(fit.mixing_probabilities, g["mixing_probabilities"]) <= 1e-13
    if fit.converged:
        assert np.array_equal(fit.labels, g["labels"])


@pytest.mark.parametrize("path", KM_FILES, ids=os.path.basename)
def test_oracle_reproduces_golden_kmeans(path):
    g = np.load(path)
    k = g["initial_means"].shape[1]
    fit = oracle.kmeans_fit(g["data"], k, init=oracle.EXPLICIT, explicit_means=g["initial_means"], maximum_steps=int(g["maximum_steps"]))
    assert fit.iterations == int(g["iterations"]) and fit.converged == bool(g["converged"])
    assert np.array_equal(fit.labels, g["labels"])
    assert abs(fit.inertia - float(g["inertia"])) <= 1e-13 * float(g["inertia"])
    assert rel(fit.centroids, g["centroids"]) <= 1e-13


@pytest.fixture(scope="module")
def ctx():
    from ml_b200 import cabi
    assert cabi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    c = cabi.Context(1)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", EM_FILES, ids=os.path.basename)
def test_cuda_reproduces_golden_em(ctx, path):
    from tests.test_gpu_cabi_parity import em_fit_cabi
    g = np.load(path)
    k = g["initial_means"].shape[1]
    fit = em_fit_cabi(ctx, g["data"], k, g["initial_means"], maximum_steps=int(g["maximum_steps"]))
    assert fit.iterations == int(g["iterations"]) and fit.converged == bool(g["converged"])
    assert abs(fit.log_likelihood - float(g["log_likelihood"])) <= RTOL * abs(float(g["log_likelihood"]))
    assert rel(fit.means, g["means"]) <= RTOL and rel(fit.covariances, g["covariances"]) <= RTOL
    assert rel(fit.mixing_probabilities, g["mixing_probabilities"]) <= RTOL
    assert np.max(np.abs(fit.responsibilities[:64] - g["responsibilities_head"])) <= 1e-9
    if fit.converged:
        assert np.array_equal(fit.labels, g["labels"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", KM_FILES, ids=os.path.basename)
def test_cuda_reproduces_golden_kmeans(ctx, path):
    from tests.test_gpu_cabi_parity import kmeans_fit_cabi
    g = np.load(path)
    k = g["initial_means"].shape[1]
    fit = kmeans_fit_cabi(ctx, g["data"], k, g["initial_means"], maximum_steps=int(g["maximum_steps"]))
    assert fit.iterations == int(g["iterations"]) and fit.converged == bool(g["converged"])
    assert np.array_equal(fit.labels, g["labels"])
    assert abs(fit.inertia - float(g["inertia"])) <= RTOL * float(g["inertia"])
    assert rel(fit.centroids, g["centroids"]) <= RTOL
