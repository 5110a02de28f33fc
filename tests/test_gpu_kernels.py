"""Unit checks of device-side building blocks through the C-ABI's diagnostic entry points."""
import numpy as np
import pytest

from ml_b200 import cabi

pytestmark = pytest.mark.gpu


def test_exp_for_nonpositive_arguments_is_accurate_to_two_ulp():
    """fastmath.cuh exp_nonpositive replaces std::exp of EM.cpp:206 after the max-shift (arguments <= 0).
    Tolerance: 2 ulp = 4.5e-16 relative against numpy's correctly-rounded-to-<1ulp exp."""
    rng = np.random.default_rng(0)
    x = np.concatenate([
        -np.abs(rng.standard_normal(200000)) * 5.0,           # the bulk of E-step arguments
        -rng.uniform(0.0, 700.0, 200000),                     # the whole supported range
        -np.exp(rng.uniform(-40.0, 6.5, 100000)),             # log-uniform magnitudes down to 4e-18
        np.array([0.0, -0.0, -1e-300, -699.999, -700.0, -700.001, -745.0, -1e4, -np.inf]),
    ])
    got = cabi.selftest_exp(x)
    want = np.exp(x)
    assert got[x.size - 9] == 1.0 and got[x.size - 8] == 1.0    # exp(0) is exactly 1
    big = x >= -700.0
    rel = np.abs(got[big] - want[big]) / want[big]
    assert rel.max() <= 4.5e-16, rel.max()
    assert np.all(got[~big] == 0.0)                            # below -700 the term is dropped (exp < 1e-304)
    assert np.all(np.diff(got[np.argsort(x)]) >= 0.0)          # monotone
