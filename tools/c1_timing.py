"""BASELINE config 1 (Benchmarks/bm_EM.cpp, bm_KMeans.cpp): mouse data, KPP, tolerance 1e-14, through the public cppyml API."""
import sys, time, json, numpy as np
sys.path.insert(0, ".")  # run from the repository root
import oracle
from ml_b200 import import_cppyml
cppyml = import_cppyml()
for n in (1000, 10000, 100000):
    data, _ = oracle.testdata_mouse(n)
    out = {"n": n}
    for rep in range(3):
        em = cppyml.clustering.EM(3); em.set_seed(42); em.set_absolute_tolerance(1e-14); em.set_relative_tolerance(1e-14)
        em.set_means_initialiser(cppyml.clustering.KPP())
        t = time.perf_counter(); em.fit(data); dt = time.perf_counter() - t
        out["em_gpu_ms"] = round(dt * 1e3, 3); out["em_iters"] = em.number_iterations
    t = time.perf_counter(); ref = oracle.em_fit(data, 3, seed=42, means_init=oracle.KPP, absolute_tolerance=1e-14, relative_tolerance=1e-14); out["em_cpu_ms"] = round((time.perf_counter() - t) * 1e3, 3); out["em_cpu_iters"] = ref.iterations
    for rep in range(3):
        km = cppyml.clustering.KMeans(3); km.set_seed(42); km.set_absolute_tolerance(1e-14); km.set_number_initialisations(3)
        km.set_centroids_initialiser(cppyml.clustering.KPP())
        t = time.perf_counter(); km.fit(data); dt = time.perf_counter() - t
        out["km_gpu_ms"] = round(dt * 1e3, 3)
    t = time.perf_counter(); kref = oracle.kmeans_fit(data, 3, seed=42, init=oracle.KPP, absolute_tolerance=1e-14, number_initialisations=3); out["km_cpu_ms"] = round((time.perf_counter() - t) * 1e3, 3)
    print(json.dumps(out), flush=True)
