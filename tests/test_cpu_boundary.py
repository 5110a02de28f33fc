"""CPU-side checks of the drop-in boundary (no GPU needed): the C-ABI library loads and exports every symbol
include/mlb200.h declares, the C++ host layer passes its own test program, the cppyml module keeps the
reference's Python surface and error behaviour, and compute entry points fail loudly without a device."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__
    __graft_entry__.build()


def header_functions():
    text = open(os.path.join(ROOT, "include", "mlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ml_b200 import cabi
    lib = cabi.lib()
    names = header_functions()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/mlb200.h but not exported by libmlb200.so"
    assert sorted(cabi.SIGNATURES) == names, "ml_b200/cabi.py and include/mlb200.h disagree on the entry points"
    assert lib.mlb_version() >= 100


def test_shard_ranges_partition_the_points():
    from ml_b200 import cabi
    for n in (1, 63, 1000, 123457, 10_000_000, 100_000_000):
        for world in (1, 2, 4, 8):
            ranges = [cabi.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            # coarser worlds are unions of the 8 virtual shards: the same cut points
            fine = [cabi.shard_range(n, 8, r) for r in range(8)]
            step = 8 // world
            assert [r[0] for r in ranges] == [fine[i * step][0] for i in range(world)]


def test_no_device_means_a_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ml_b200 import cabi
    with pytest.raises(cabi.MlbError) as err:
        cabi.Context(1)
    assert err.value.code == cabi.MLB_ECUDA
    cppyml = __import__("ml_b200").import_cppyml()
    em = cppyml.clustering.EM(2)
    with pytest.raises(RuntimeError):
        em.fit(np.random.default_rng(0).normal(size=(50, 2)))
    km = cppyml.clustering.KMeans(2)
    with pytest.raises(RuntimeError):
        km.fit(np.random.default_rng(0).normal(size=(50, 2)))


def test_invalid_arguments_are_rejected_before_any_device_work():
    from ml_b200 import cabi
    lib = cabi.lib()
    import ctypes
    h = ctypes.c_void_p()
    assert lib.mlb_ctx_create(None, 3, ctypes.byref(h)) == cabi.MLB_EINVAL      # 3 GPUs: not a divisor of the 8 virtual shards
    assert b"1, 2, 4 or 8" in lib.mlb_last_error()
    b, e = ctypes.c_int64(), ctypes.c_int64()
    assert lib.mlb_shard_range(100, 3, 0, ctypes.byref(b), ctypes.byref(e)) == cabi.MLB_EINVAL
    assert lib.mlb_em_step(None, None) == cabi.MLB_EINVAL
    assert lib.mlb_data_kpp_update(None, None, 1, None) == cabi.MLB_EINVAL
    assert lib.mlb_em_mstep_from_labels(None, None) == cabi.MLB_EINVAL
    assert lib.mlb_em_predict(None, None, 0, 0, None, 0, None) == cabi.MLB_EINVAL
    assert lib.mlb_km_predict(None, None, 0, 0, None, None) == cabi.MLB_EINVAL


def test_host_layer_cpp_tests():
    """tests/cpp/host_tests.cpp: initialiser PRNG streams bit-equal to the oracle, the streamed discrete draw of the
    device-assisted KPP bit-equal to std::discrete_distribution, LinearAlgebra identities,
    argument errors, N == K exact fits (Tests/test_EM.cpp:126-144, Tests/test_KMeans.cpp:108-127)."""
    exe = os.path.join(ROOT, "build", "host_tests")
    lib_dir, oracle_dir = os.path.join(ROOT, "ml_b200", "lib"), os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "ml_b200", "host"), "-I", os.path.join(ROOT, "include"),
                           "-isystem", os.path.join(ROOT, "ml_b200", "host", "eigen_compat"), os.path.join(ROOT, "tests", "cpp", "host_tests.cpp"),
                           "-o", exe, "-L", lib_dir, "-lML", "-lmlb200", "-L", oracle_dir, "-lmlpp_oracle",
                           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{oracle_dir}"])
    out = subprocess.check_output([exe], text=True)
    assert out.startswith("ok ")


def test_cppyml_surface_and_errors():
    """The names and error mapping of cppyml/clustering.cpp:75-184 and cppyml/tests/test_clustering.py."""
    cppyml = __import__("ml_b200").import_cppyml()
    c = cppyml.clustering
    for name in ("CentroidsInitialiser", "ResponsibilitiesInitialiser", "Forgy", "RandomPartition", "KPP", "ClosestCentroid", "EM", "KMeans"):
        assert hasattr(c, name)
    assert issubclass(c.KPP, c.CentroidsInitialiser) and issubclass(c.ClosestCentroid, c.ResponsibilitiesInitialiser)
    em = c.EM(3)
    for method in ("set_seed", "set_absolute_tolerance", "set_relative_tolerance", "set_maximum_steps", "set_means_initialiser",
                   "set_responsibilities_initialiser", "set_verbose", "set_maximise_first", "fit", "covariance", "assign_responsibilities"):
        assert callable(getattr(em, method))
    assert em.number_components == 3
    em.set_means_initialiser(c.KPP())
    em.set_responsibilities_initialiser(c.ClosestCentroid(c.RandomPartition()))
    with pytest.raises(ValueError):
        c.EM(0)
    with pytest.raises(ValueError):
        em.set_absolute_tolerance(-1.0)
    with pytest.raises(ValueError):
        em.set_relative_tolerance(-1.0)
    with pytest.raises(ValueError):
        em.set_maximum_steps(1)
    with pytest.raises(ValueError):
        em.fit(np.zeros((2, 2)))                      # fewer points than components
    with pytest.raises(TypeError):
        em.fit(np.zeros((10, 2), dtype=np.float32))   # noconvert: wrong dtype is an error, never a silent copy
    with pytest.raises(TypeError):
        em.fit(np.asfortranarray(np.zeros((10, 2))))  # wrong layout likewise
    with pytest.raises(TypeError):
        em.fit([[0.0, 1.0]] * 10)
    km = c.KMeans(3)
    for method in ("set_seed", "set_absolute_tolerance", "set_maximum_steps", "set_centroids_initialiser", "set_number_initialisations",
                   "set_verbose", "fit", "assign_label"):
        assert callable(getattr(km, method))
    with pytest.raises(ValueError):
        c.KMeans(0)
    with pytest.raises(ValueError):
        km.set_number_initialisations(0)
    # N == K: the exact fits need no device
    x = np.array([[0.5, 0.1], [0.3, 0.2], [0.1, -0.4]])
    assert em.fit(x) is True
    assert em.means.shape == (2, 3) and np.array_equal(em.means, x.T)
    assert em.responsibilities.shape == (3, 3) and np.array_equal(em.responsibilities, np.eye(3))
    assert em.log_likelihood == np.inf and em.mixing_probabilities.shape == (3,)
    assert km.fit(x) is True
    assert km.centroids.shape == (3, 2) and np.array_equal(km.centroids, x)
    assert km.labels == [0, 1, 2] and km.inertia == 0.0
    assert km.assign_label(x[1]) == (1, 0.0)
    # the batched forms run on the device only: wrong shapes are ValueErrors, a model without device state fails loudly
    assert callable(em.assign_responsibilities_batch) and callable(km.assign_labels)
    with pytest.raises(ValueError):
        em.assign_responsibilities_batch(np.zeros((4, 5)))
    with pytest.raises(ValueError):
        km.assign_labels(np.zeros((4, 5)))
    with pytest.raises(RuntimeError):
        em.assign_responsibilities_batch(np.zeros((4, 2)))   # N == K exact fit: nothing is resident
    with pytest.raises(RuntimeError):
        km.assign_labels(np.zeros((4, 2)))
    with pytest.raises(TypeError):
        km.assign_labels(np.zeros((4, 2), dtype=np.float32))


def test_bench_reference_arm_runs_on_cpu():
    import json
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                                   "--cpu-sample", "20000"], text=True, cwd=ROOT)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Gpoint*comp/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "ml_b200")):
        for name in files:
            if name.endswith((".py", ".cpp", ".hpp", ".cu", ".h", "Makefile")):
                text = open(os.path.join(base, name), errors="replace").read()
                assert "mlpp_oracle" not in text and "import oracle" not in text, os.path.join(base, name)
