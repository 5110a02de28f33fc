"""Multi-GPU parity on real devices (skipped with fewer than 2): the same fit on G = 1 and G = 2 (4, 8 when present)
GPUs of one process must agree BITWISE (fixed chunks, 8 virtual shards, one all-gather, fixed-tree sum), and the
torchrun-style rank contexts must agree with the single-process context."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.datasets import synthetic_gmm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from ml_b200 import cabi
    return cabi.device_count()


def _em_run(n_devices, data, k, steps):
    from ml_b200 import cabi
    ctx = cabi.Context(n_devices)
    d_data = cabi.Data.upload(ctx, data)
    em = cabi.Em(d_data, k)
    cov = em.sample_covariance()
    em.set_params(np.ascontiguousarray(data[:k].T), np.repeat(cov[None], k, axis=0), np.full(k, 1.0 / k))
    lls = [em.step() for _ in range(steps)]
    means, covs, w = em.get_params()
    resp, labels = em.emit()
    em.close(); d_data.close(); ctx.close()
    return np.array(lls), means, covs, w, resp, labels, cov


def _km_run(n_devices, data, k, steps):
    from ml_b200 import cabi
    ctx = cabi.Context(n_devices)
    d_data = cabi.Data.upload(ctx, data)
    km = cabi.Km(d_data, k)
    km.set_centroids(np.ascontiguousarray(data[:k].T))
    trace = []
    for _ in range(steps):
        inertia, changed = km.assign()
        trace.append((inertia, changed, km.update()))
    cents, labels = km.get_centroids(), km.get_labels()
    km.close(); d_data.close(); ctx.close()
    return np.array(trace), cents, labels


@pytest.mark.parametrize("n,d,k", [(40013, 8, 16), (30000, 16, 32), (20000, 24, 40)])
def test_em_is_bitwise_invariant_in_the_gpu_count(n, d, k):
    have = _device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    data, _, _ = synthetic_gmm(n, d, k, seed=61)
    ref = _em_run(1, data, k, 4)
    for g in (2, 4, 8):
        if g > have:
            break
        got = _em_run(g, data, k, 4)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), g


def test_kmeans_is_bitwise_invariant_in_the_gpu_count():
    have = _device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    data, _, _ = synthetic_gmm(50021, 16, 20, seed=62)
    ref = _km_run(1, data, 48, 4)
    for g in (2, 4, 8):
        if g > have:
            break
        got = _km_run(g, data, 48, 4)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), g


def _kms_run(n_devices, data, k, starts):
    from ml_b200 import cabi
    ctx = cabi.Context(n_devices)
    d_data = cabi.Data.upload(ctx, data)
    sets = cabi.Kms(d_data, k, len(starts))
    for s, c in enumerate(starts):
        sets.set_centroids(s, c)
    trace = []
    for _ in range(3):
        inertia, changed = sets.assign()
        trace.append(np.concatenate([inertia, changed.astype(np.float64), sets.update()]))
    inertia, changed = sets.assign(0b101)
    out = [np.array(trace), inertia.copy(), changed.copy()] + [sets.get_labels(s) for s in range(len(starts))] + [sets.get_centroids(s) for s in range(len(starts))]
    sets.close(); d_data.close(); ctx.close()
    return out


@pytest.mark.parametrize("n,d,k", [(30011, 8, 20), (20000, 16, 48)])
def test_kmeans_start_sets_are_bitwise_invariant_in_the_gpu_count(n, d, k):
    have = _device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    data, _, _ = synthetic_gmm(n, d, 10, seed=64)
    rng = np.random.default_rng(9)
    starts = [np.ascontiguousarray(data[rng.choice(n, size=k, replace=False)].T) for _ in range(3)]
    ref = _kms_run(1, data, k, starts)
    for g in (2, 4, 8):
        if g > have:
            break
        got = _kms_run(g, data, k, starts)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), g


def _seeding_run(n_devices, data, k, labels):
    """KPP distance passes, the M-step from labels and prediction on a G-GPU context."""
    from ml_b200 import cabi
    ctx = cabi.Context(n_devices)
    d_data = cabi.Data.upload(ctx, data)
    nearest = [d_data.kpp_update(data[i * 7], first=(i == 0)).copy() for i in range(3)]
    em = cabi.Em(d_data, k)
    em.mstep_from_labels(labels)
    params = em.get_params()
    resp, pred = em.predict(data[:3000])
    km = cabi.Km(d_data, k)
    km.set_centroids(np.ascontiguousarray(data[:k].T))
    km_labels, km_dist = km.predict(data[:3000])
    km.close(); em.close(); d_data.close(); ctx.close()
    return (*nearest, *params, resp, pred, km_labels, km_dist)


def test_seeding_and_prediction_are_bitwise_invariant_in_the_gpu_count():
    have = _device_count()
    if have < 2:
        pytest.skip("needs at least 2 GPUs")
    data, labels, _ = synthetic_gmm(30011, 8, 16, seed=63)
    ref = _seeding_run(1, data, 16, labels)
    for g in (2, 4, 8):
        if g > have:
            break
        got = _seeding_run(g, data, 16, labels)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), g


def test_bench_under_torchrun_two_ranks():
    """bench.py launched the way the driver launches it for N = 2: one rank per GPU, one JSON line from rank 0."""
    if _device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["gpu_launches"] > 0
