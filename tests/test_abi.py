"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/tloam_b200.h
declares, the POD config mirrors the YAML defaults, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tloam_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tloam_b200.h")).read()
    return sorted(set(re.findall(r"\b(tloam_b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tloam_b200.h but not exported"


def test_binding_lists_the_same_symbols():
    from tloam_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_default_config_matches_reference_yaml(lib):
    """ref: config/mapping/lidar_odometry.yaml:23-39"""
    from tloam_b200 import default_config
    c = default_config()
    assert (c.k_corr, c.factor_num) == (10, 4)
    assert (c.edge_dist_thres, c.sphere_dist_thres, c.planar_dist_thres, c.ground_dist_thres) == (1.0, 0.5, 0.5, 0.5)
    assert c.edge_dir_thres == 0.85
    assert (c.edge_maxnum, c.sphere_maxnum, c.planar_maxnum, c.ground_maxnum) == (1200, 200, 2500, 2000)
    assert (c.max_iterations, c.cost_threshold, c.gnc_factor, c.noise_bound, c.fitness_thres) == (4, 5e-9, 11.8, 0.01, 0.02)
    assert c.ceres_max_num_iterations == 4


def test_config_struct_layout_matches_oracle_prefix(lib, oracle):
    """The oracle's config carries the same 16 YAML fields in the same order (then its own thread switches)."""
    from tloam_b200 import _lib
    a = [f[0] for f in _lib.TlsConfig._fields_]
    b = [f[0] for f in oracle.Config._fields_]
    assert b[:len(a)] == a


def test_status_strings(lib):
    assert lib.tloam_b200_status_string(0) == b"ok"
    assert b"no CPU fallback" in lib.tloam_b200_status_string(5)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tloam_b200
    with pytest.raises(tloam_b200.RegistrationError) as e:
        tloam_b200.LocalRegistration()
    assert e.value.status == 5


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tloam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "tloam_oracle" not in text and "from oracle" not in text, f
