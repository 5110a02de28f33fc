"""Generates the golden fixtures of tests/golden/ from the REFERENCE'S OWN CODE (oracle/_ref: ML/EM.cpp, ML/KMeans.cpp,
ML/Clustering.cpp, ML/LinearAlgebra.cpp compiled where they lie under /root/reference against the Eigen stand-in of
oracle/eigen_standin/).  The reference ships no numeric golden vectors of its own, so these are outputs
of the reference run in the builder container, committed so that the GPU box (which has no /root/reference) can check
both the oracle port and the CUDA path against them.

    python tests/golden/make_golden.py        # needs oracle/_ref/libmlpp_ref.so

Each fixture stores the input points (small), the explicit initial means, the options, and the reference's results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tests.datasets import synthetic_gmm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

EM_CASES = [  # name, n, d, k, seed, maximum_steps
    ("em_d2_k3", 1500, 2, 3, 101, 200),
    ("em_d8_k16", 3000, 8, 16, 102, 60),
    ("em_d16_k32", 4000, 16, 32, 103, 25),
    ("em_d5_k7", 1200, 5, 7, 104, 200),
    ("em_d24_k6", 2000, 24, 6, 105, 30),     # split E / M kernels
    ("em_d6_k40", 8000, 6, 40, 106, 20),     # split path, K > 32 (200 points per component: a well-posed fit)
]
KM_CASES = [
    ("km_d2_k3", 1500, 2, 3, 201),
    ("km_d8_k16", 3000, 8, 16, 202),
    ("km_d32_k64", 3000, 32, 64, 203),
    ("km_d16_k33", 2000, 16, 33, 204),
]


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref, needs /root/reference)"
    for name, n, d, k, seed, steps in EM_CASES:
        data, _, _ = synthetic_gmm(n, d, k, seed=seed, spread=6.0)
        init = np.ascontiguousarray(data[3:: n // k][:k].T)
        fit = oracle.em_fit(data, k, means_init=oracle.EXPLICIT, explicit_means=init, maximum_steps=steps, impl="reference")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=data, initial_means=init, maximum_steps=steps,
                            absolute_tolerance=1e-8, relative_tolerance=1e-8, converged=fit.converged, iterations=fit.iterations,
                            log_likelihood=fit.log_likelihood, means=fit.means, covariances=fit.covariances,
                            mixing_probabilities=fit.mixing_probabilities, labels=fit.labels,
                            responsibilities_head=fit.responsibilities[:64])
        print(name, "iterations", fit.iterations, "converged", fit.converged, "ll", fit.log_likelihood)
    for name, n, d, k, seed in KM_CASES:
        data, _, _ = synthetic_gmm(n, d, min(k, 20), seed=seed, spread=5.0)
        init = np.ascontiguousarray(data[5:: n // k][:k].T)
        fit = oracle.kmeans_fit(data, k, init=oracle.EXPLICIT, explicit_means=init, maximum_steps=200, impl="reference")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=data, initial_means=init, maximum_steps=200, absolute_tolerance=1e-8,
                            converged=fit.converged, iterations=fit.iterations, inertia=fit.inertia, centroids=fit.centroids, labels=fit.labels)
        print(name, "iterations", fit.iterations, "converged", fit.converged, "inertia", fit.inertia)


if __name__ == "__main__":
    main()
