"""The CUDA path at BASELINE.json's full single-GPU sizes (C2: N = 10M, D = 8, K = 16; the per-GPU shares of C3:
N = 12.5M, D = 16, K = 32 and C5: N = 12.5M, D = 32, K = 256), where the single-threaded oracle would need minutes to
hours per iteration.  Checked instead through properties that do not depend on the size:

  * EM never decreases the log-likelihood (Dempster, Laird, Rubin), mixing weights sum to one, covariances are
    symmetric positive definite;
  * one whole iteration recomputed with vectorised numpy (an independent formulation: Cholesky solves and a
    log-sum-exp, pairwise sums) from the same parameters gives the same log-likelihood, weights, means and
    covariances within 1e-9 relative (the north star's tolerance), the same labels and the same responsibilities
    on sampled ranges, including the ragged tail;
  * Lloyd iterations never increase the inertia; the inertia is the sum of the squared distances to the assigned
    centroids; every label is the arg-min of the directly evaluated distances (sampled); every updated centroid is
    the mean of its points.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


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


def _download_all(data, n, block=1 << 22):
    d = data.shape[2]
    out = np.empty((n, d))
    for lo in range(0, n, block):
        out[lo:lo + block] = data.download(lo, min(block, n - lo))
    return out


def _numpy_em_iteration(x, means_dk, covs, weights, block=1 << 17):
    """One E-step + M-step of EM.cpp:190-263 in blocks of rows, as a few large matrix products: returns (mean
    log-likelihood, weights, means (D, K), covariances (K, D, D), labels, responsibilities of the first and last 1000 rows).
    y_k = L_k^-1 (x - mu_k) for all components at once is one product with the stacked inverse factors; the second
    moments about a nearby point are one product of the pairwise products z_a z_b with the responsibilities."""
    n, d = x.shape
    k = weights.size
    chol = [np.linalg.cholesky(covs[c]) for c in range(k)]
    log_norm = np.array([np.log(weights[c]) - np.log(np.diag(chol[c])).sum() for c in range(k)])
    inv_l = [np.linalg.inv(chol[c]) for c in range(k)]
    stacked = np.concatenate([inv_l[c].T for c in range(k)], axis=1)                     # D x (K D)
    offsets = np.concatenate([means_dk[:, c] @ inv_l[c].T for c in range(k)])            # K D
    iu = np.triu_indices(d)
    s0 = np.zeros(k)
    s1 = np.zeros((k, d))
    s2u = np.zeros((iu[0].size, k))
    ll = 0.0
    labels = np.empty(n, dtype=np.uint32)
    head = tail = None
    centre = x[: 1 << 16].mean(axis=0)   # moments about a nearby point: no cancellation in the second moment
    for lo in range(0, n, block):
        xb = x[lo:lo + block]
        y = xb @ stacked
        y -= offsets
        np.square(y, out=y)
        logp = log_norm - 0.5 * y.reshape(xb.shape[0], k, d).sum(axis=2)
        mx = logp.max(axis=1, keepdims=True)
        e = np.exp(logp - mx)
        tot = e.sum(axis=1, keepdims=True)
        ll += float((mx[:, 0] + np.log(tot[:, 0])).sum())
        r = e / tot
        labels[lo:lo + block] = np.argmax(r, axis=1)
        if lo == 0:
            head = r[:1000].copy()
        if lo + block >= n:
            tail = r[-1000:].copy()
        z = xb - centre
        s0 += r.sum(axis=0)
        s1 += r.T @ z
        s2u += (z[:, iu[0]] * z[:, iu[1]]).T @ r
    new_w = s0 / n
    m1 = s1 / s0[:, None]
    new_means = (centre + m1).T
    new_covs = np.empty((k, d, d))
    for c in range(k):
        m2 = np.zeros((d, d))
        m2[iu] = s2u[:, c]
        m2 = m2 + m2.T - np.diag(np.diag(m2))
        new_covs[c] = m2 / s0[c] - np.outer(m1[c], m1[c]) + 1e-15 * np.eye(d)
    return ll / n - 0.5 * d * np.log(2 * np.pi), new_w, new_means, new_covs, labels, head, tail


def _start(cabi, ctx, n, d, k):
    data = cabi.Data.generate_gmm(ctx, n, d, k, seed=20261018)
    em = cabi.Em(data, k)
    init = data.download(0, k).T
    em.set_params(init, np.repeat(em.sample_covariance()[None], k, axis=0), np.full(k, 1.0 / k))
    lls = em.run_steps(6)
    # monotone log-likelihood (rounding: the sum over 1e7 points is good to ~1e-13 relative)
    assert np.all(np.diff(lls) >= -1e-11 * np.abs(lls[:-1])), lls
    before = em.get_params()
    ll_next = em.step()
    assert ll_next >= lls[-1] - 1e-11 * abs(lls[-1])
    means, covs, weights = em.get_params()
    assert abs(weights.sum() - 1.0) <= 1e-12
    for c in range(k):
        assert np.array_equal(covs[c], covs[c].T)
        np.linalg.cholesky(covs[c])
    return data, em, before, ll_next, (means, covs, weights)


def test_em_c2_at_full_size_against_a_numpy_iteration(cabi, ctx):
    """C2 as quoted: N = 10M, D = 8, K = 16 on one GPU; the whole iteration recomputed on the host."""
    n, d, k = 10_000_000, 8, 16
    data, em, before, ll_next, (means, covs, weights) = _start(cabi, ctx, n, d, k)
    x = _download_all(data, n)
    want_ll, want_w, want_means, want_covs, want_labels, head, tail = _numpy_em_iteration(x, *before)
    assert abs(ll_next - want_ll) <= RTOL * abs(want_ll)
    assert np.max(np.abs(weights - want_w)) <= RTOL * np.max(want_w)
    assert np.max(np.abs(means - want_means)) <= RTOL * np.max(np.abs(want_means))
    for c in range(k):
        assert np.max(np.abs(covs[c] - want_covs[c])) <= RTOL * np.max(np.abs(want_covs[c])), c
    # what the reference leaves behind after that step: responsibilities and labels at the parameters its E-step used
    got_head, _ = em.emit_range(0, 1000, want_labels=False)
    got_tail, _ = em.emit_range(n - 1000, 1000, want_labels=False)
    assert np.max(np.abs(got_head - head)) <= 1e-9 and np.max(np.abs(got_tail - tail)) <= 1e-9
    _, labels = em.emit(want_responsibilities=False)
    differ = np.flatnonzero(labels != want_labels)
    assert differ.size <= n * 1e-6   # only rows whose two largest responsibilities agree to rounding may differ
    em.close(), data.close()


@pytest.mark.parametrize("gpus", [1, 8])
def test_em_c3_at_full_size_through_moment_identities(cabi, gpus):
    """C3 (D = 16, K = 32) with 12.5M points per GPU: the per-GPU share on one GPU, and the north-star configuration as
    quoted, N = 100M point-sharded over 8 GPUs of one process, when the box has them.  Every row of responsibilities
    sums to one, so the M-step's outputs must recombine to plain sums over the data, whatever the responsibilities are:
        sum_k s_k = N,   sum_k s_k mu_k = sum_i x_i,   sum_k s_k (Sigma_k - 1e-15 I + mu_k mu_k^T) = sum_i x_i x_i^T,
    which checks that every point of every shard (ragged tail included) entered the statistics exactly once; sampled
    rows of the responsibilities are recomputed on the host."""
    if cabi.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    n, d, k = 12_500_000 * gpus, 16, 32
    ctx = cabi.Context(gpus)
    data, em, before, ll_next, (means, covs, weights) = _start(cabi, ctx, n, d, k)
    x = _download_all(data, n)
    centre = x[: 1 << 16].mean(axis=0)
    first = np.zeros(d)
    second = np.zeros((d, d))
    for lo in range(0, n, 1 << 22):
        z = x[lo:lo + (1 << 22)] - centre
        first += z.sum(axis=0)
        second += z.T @ z
    first /= n
    second /= n
    mu = means.T - centre                                        # (K, D), about the same point
    got_first = weights @ mu
    got_second = sum(weights[c] * (covs[c] - 1e-15 * np.eye(d) + np.outer(mu[c], mu[c])) for c in range(k))
    assert np.max(np.abs(got_first - first)) <= 1e-10 * np.sqrt(np.max(np.diag(second)))
    assert np.max(np.abs(got_second - second)) <= 1e-10 * np.max(np.abs(second))
    rows = np.concatenate([np.arange(1000), np.arange(n - 1000, n)])
    _, _, _, _, _, head, tail = _numpy_em_iteration(np.ascontiguousarray(x[rows]), *before)
    got_head, lab_head = em.emit_range(0, 1000)
    got_tail, lab_tail = em.emit_range(n - 1000, 1000)
    assert np.max(np.abs(got_head - head)) <= 1e-9 and np.max(np.abs(got_tail - tail)) <= 1e-9
    clear = np.sort(head, axis=1)[:, -1] - np.sort(head, axis=1)[:, -2] > 1e-6
    assert np.array_equal(lab_head[clear], np.argmax(head, axis=1)[clear].astype(np.uint32))
    em.close(), data.close(), ctx.close()


def test_kmeans_at_full_size(cabi, ctx):
    n, d, k = 12_500_000, 32, 256
    data = cabi.Data.generate_gmm(ctx, n, d, 64, seed=20261018)
    km = cabi.Km(data, k)
    km.set_centroids(data.download(0, k).T)
    inertias = []
    for _ in range(5):
        inertia, changed = km.assign()
        inertias.append(inertia)
        km.update()
    assert np.all(np.diff(inertias) <= 1e-12 * np.array(inertias[:-1])), inertias   # Lloyd never increases the inertia
    centroids = km.get_centroids()                     # (D, K), the ones the next assignment uses
    inertia, changed = km.assign()
    labels = km.get_labels()
    assert 0 <= changed <= n and labels.max() < k
    x = _download_all(data, n)
    # inertia = sum of squared distances to the assigned centroids (KMeans.cpp:167-178), recomputed directly
    total = 0.0
    for lo in range(0, n, 1 << 21):
        diff = x[lo:lo + (1 << 21)] - centroids.T[labels[lo:lo + (1 << 21)]]
        total += float(np.einsum("ij,ij->", diff, diff))
    assert abs(inertia - total) <= 1e-10 * total
    # labels are the arg-min of the directly evaluated distances: a sample of rows, head, tail and random
    rng = np.random.default_rng(3)
    rows = np.unique(np.concatenate([np.arange(2000), np.arange(n - 2000, n), rng.integers(0, n, size=4000)]))
    d2 = ((x[rows][:, None, :] - centroids.T[None]) ** 2).sum(axis=2)
    best = d2.argmin(axis=1)
    gap = np.partition(d2, 1, axis=1)
    clear = gap[:, 1] - gap[:, 0] > 1e-9 * gap[:, 1]
    assert np.array_equal(labels[rows][clear], best[clear].astype(np.uint32))
    # update_step (KMeans.cpp:180-192): every centroid becomes the mean of its points
    km.update()
    updated = km.get_centroids()
    counts = np.bincount(labels, minlength=k).astype(np.float64)
    sums = np.stack([np.bincount(labels, weights=x[:, j], minlength=k) for j in range(d)])
    want = np.where(counts > 0, sums / np.maximum(counts, 1.0), 0.0)
    assert np.max(np.abs(updated - want)) <= 1e-10 * np.max(np.abs(want))
    km.close(), data.close()
