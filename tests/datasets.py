"""Data sets shared by the tests.  Each one restates the generator of a reference test or benchmark."""
import numpy as np


def mouse_numpy(sample_size=1000, seed=999):
    """The "mouse" data of cppyml/tests/test_clustering.py:19-45 (three uniform discs, D=2).

    Same numpy calls in the same order, so the sample equals the reference test's sample.
    """
    np.random.seed(seed)
    dims = 2
    face_radius = 1
    ear_radius = 0.3
    ear_angle = np.pi / 4
    num_components = 3
    radii = (face_radius, ear_radius, ear_radius)
    ear_weight = 2
    weights = [face_radius ** 2, ear_weight * ear_radius ** 2, ear_weight * ear_radius ** 2]
    probabilities = weights / np.sum(weights)
    indices = np.random.choice(np.arange(num_components), sample_size, p=probabilities)
    center_xs = [0, (face_radius + ear_radius) * np.sin(-ear_angle), (face_radius + ear_radius) * np.sin(ear_angle)]
    center_ys = [0, (face_radius + ear_radius) * np.cos(-ear_angle), (face_radius + ear_radius) * np.cos(ear_angle)]
    data = np.empty((sample_size, dims))
    for i in range(sample_size):
        k = indices[i]
        phi = np.random.rand() * 2 * np.pi
        r = np.sqrt(np.random.rand()) * radii[k]
        data[i, 0] = center_xs[k] + r * np.cos(phi)
        data[i, 1] = center_ys[k] + r * np.sin(phi)
    return data


def synthetic_gmm(n, d, k, seed=0, spread=10.0):
    """Numpy twin of the bench generator's *distribution* (SURVEY.md §8d): K means U[-spread, spread]^D,
    covariances A A^T / D + 0.5 I, Dirichlet(5) weights.  Returns (data (N, D), true labels, true means (K, D))."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-spread, spread, size=(k, d))
    chols = []
    for _ in range(k):
        a = rng.standard_normal((d, d))
        chols.append(np.linalg.cholesky(a @ a.T / d + 0.5 * np.eye(d)))
    weights = rng.dirichlet(np.full(k, 5.0))
    labels = rng.choice(k, size=n, p=weights)
    z = rng.standard_normal((n, d))
    data = np.empty((n, d))
    for c in range(k):
        m = labels == c
        data[m] = means[c] + z[m] @ chols[c].T
    return np.ascontiguousarray(data), labels.astype(np.uint32), means


def separated_clusters(n, d, k, separation, offset=0.0, seed=0):
    """K unit-variance Gaussian clusters whose centres lie `separation` standard deviations apart along a line through
    the D-cube's diagonal, all shifted by `offset` in every coordinate: the regime of un-standardised features, where the
    distance of a component from the centre of the data, measured in its own standard deviations, is large.  Returns
    (data (N, D), true labels, centres (K, D))."""
    rng = np.random.default_rng(seed)
    direction = np.ones(d) / np.sqrt(d)
    centres = np.array([(j - (k - 1) / 2.0) * separation * direction for j in range(k)])
    centres += rng.uniform(-0.25, 0.25, size=(k, d)) * separation   # not collinear
    centres += offset
    labels = rng.integers(0, k, size=n)
    scales = rng.uniform(0.7, 1.4, size=(k, d))
    data = centres[labels] + rng.standard_normal((n, d)) * scales[labels]
    return np.ascontiguousarray(data), labels.astype(np.uint32), centres
