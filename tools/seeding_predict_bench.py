"""Timing of the SURVEY.md §8(f) rows (not the bench.py contract): the K-means++ distance pass, the device-assisted
KPP initialisation against the CPU oracle's, and batched prediction.  One JSON line per case.

  python tools/seeding_predict_bench.py            # all cases, sizes that finish in about a minute on one B200
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ml_b200 import cabi, import_cppyml

HBM_PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def kpp_pass_case(ctx, n, d, passes=10):
    """mlb_data_kpp_update alone, data resident, nothing downloaded: algorithmic bytes = 8 D (points) + 16 (nearest read
    and written) per point; the first pass does not read `nearest`."""
    data = cabi.Data.generate_gmm(ctx, n, d, 8, seed=3)
    c = data.download(0, passes + 3)
    for i in range(3):
        data.kpp_update(c[i], first=(i == 0), want_nearest=False)
    ctx.timer_start()
    for i in range(passes):
        data.kpp_update(c[3 + i], first=False, want_nearest=False)
    ms = ctx.timer_stop() / passes
    t0 = time.perf_counter()
    nearest = data.kpp_update(c[0], first=False, want_nearest=True)
    with_download_ms = (time.perf_counter() - t0) * 1e3
    gbs = (8 * d + 16) * n / ms / 1e6
    data.close()
    return dict(kind="kpp_pass", n=n, d=d, ms_per_pass=ms, hbm_gbs=gbs, hbm_frac=gbs / HBM_PEAK, hbm_peak_gbs=HBM_PEAK,
                pass_plus_download_ms=with_download_ms, nearest_min=float(nearest.min()))


def kpp_init_case(n, d, k):
    """The whole initialisation through the public API: cppyml KMeans with the KPP initialiser and 2 steps, against
    the oracle's KPP (the reference's O(N K^2 D) loop, one core) on the same data and seed."""
    import oracle   # CPU baseline only
    from tests.datasets import synthetic_gmm
    clustering = import_cppyml().clustering
    data, _, _ = synthetic_gmm(n, d, k, seed=4, spread=8.0)
    km = clustering.KMeans(k)
    km.set_maximum_steps(2)
    km.fit(data)                                # warm-up: context creation, first allocations
    km.set_seed(17)
    km.set_centroids_initialiser(clustering.Forgy())
    t0 = time.perf_counter()
    km.fit(data)
    forgy_s = time.perf_counter() - t0          # upload + Forgy + 2 Lloyd steps
    km.set_seed(17)
    km.set_centroids_initialiser(clustering.KPP())
    t0 = time.perf_counter()
    km.fit(data)
    kpp_s = time.perf_counter() - t0            # upload + device-assisted KPP + 2 Lloyd steps
    t0 = time.perf_counter()
    ref = oracle.centroids_init(oracle.KPP, data, k, seed=17)
    cpu_s = time.perf_counter() - t0
    return dict(kind="kpp_init", n=n, d=d, k=k, fit2_forgy_s=forgy_s, fit2_kpp_s=kpp_s, kpp_device_assisted_s=kpp_s - forgy_s,
                kpp_oracle_1core_s=cpu_s, speedup=cpu_s / max(kpp_s - forgy_s, 1e-9), oracle_centroids_finite=bool(np.isfinite(ref).all()))


def em_predict_case(ctx, n_fit, m, d, k):
    data = cabi.Data.generate_gmm(ctx, n_fit, d, k, seed=7)
    em = cabi.Em(data, k)
    em.set_params(data.download(0, k).T, np.repeat(em.sample_covariance()[None], k, axis=0), np.full(k, 1.0 / k))
    em.run_steps(3)
    queries = np.ascontiguousarray(np.tile(data.download(0, n_fit), (m // n_fit, 1)))
    em.predict(queries[:1000])
    t0 = time.perf_counter()
    _, labels = em.predict(queries, want_responsibilities=False)
    labels_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    resp, _ = em.predict(queries, want_labels=False)
    resp_s = time.perf_counter() - t0
    em.close(); data.close()
    return dict(kind="em_predict", m=m, d=d, k=k, labels_only_s=labels_s, labels_only_mpoints_per_s=m / labels_s / 1e6,
                responsibilities_s=resp_s, responsibilities_mpoints_per_s=m / resp_s / 1e6,
                h2d_bytes=8 * d * m, d2h_bytes_resp=8 * k * m, row_sum_err=float(np.max(np.abs(resp.sum(axis=1) - 1))))


def km_predict_case(ctx, n_fit, m, d, k):
    data = cabi.Data.generate_gmm(ctx, n_fit, d, min(k, 64), seed=9)
    km = cabi.Km(data, k)
    km.set_centroids(data.download(0, k).T)
    km.assign(); km.update()
    queries = np.ascontiguousarray(np.tile(data.download(0, n_fit), (m // n_fit, 1)))
    km.predict(queries[:1000])
    t0 = time.perf_counter()
    labels, dist = km.predict(queries)
    s = time.perf_counter() - t0
    km.close(); data.close()
    return dict(kind="km_predict", m=m, d=d, k=k, seconds=s, mpoints_per_s=m / s / 1e6, gpc_per_s=m * k / s / 1e9,
                h2d_bytes=8 * d * m, d2h_bytes=12 * m, inertia=float(dist.sum()))


if __name__ == "__main__":
    only = set(sys.argv[1:])   # e.g. `kpp_init`; default: everything
    want = lambda kind: not only or kind in only
    ctx = cabi.Context(1)
    if want("kpp_pass"):
        for n, d in [(12_500_000, 16), (12_500_000, 32), (10_000_000, 8), (2_500_000, 64)]:
            print(json.dumps(kpp_pass_case(ctx, n, d)), flush=True)
    if want("em_predict"):
        print(json.dumps(em_predict_case(ctx, 1_000_000, 4_000_000, 16, 32)), flush=True)
        print(json.dumps(em_predict_case(ctx, 1_000_000, 4_000_000, 8, 16)), flush=True)
    if want("km_predict"):
        print(json.dumps(km_predict_case(ctx, 1_000_000, 4_000_000, 32, 256)), flush=True)
    ctx.close()
    if want("kpp_init"):
        print(json.dumps(kpp_init_case(2_000_000, 16, 32)), flush=True)
