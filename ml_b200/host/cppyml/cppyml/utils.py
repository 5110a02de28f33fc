"""Utility functions of the package (the reference's cppyml/cppyml/utils.py), computed on the GPU."""
import numpy as np

from .cppyml import _standardise_features


def standardise_features(features: np.ndarray) -> np.ndarray:
    """Standardises an N x D feature matrix with data points in rows.

    Every column has its mean subtracted and, when N > 1, is divided by its *biased* standard deviation
    (same contract as the reference's function; the two passes over the data run on the device).

    Args:
        features: N x D feature matrix.

    Returns:
        A standardised copy.
    """
    if len(features.shape) != 2:
        raise ValueError(f"Features matrix must be 2D, got {features.shape}")
    if not features.size:
        return features.copy()
    x = np.ascontiguousarray(features, dtype=np.float64)
    return _standardise_features(x)
