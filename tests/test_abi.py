"""C-ABI surface: the library loads and exports every symbol include/dekf_b200.h declares; without a
CUDA device dekf_create refuses (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from decentralized_ekf_mhe_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dekf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dekf_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from decentralized_ekf_mhe_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_config_struct_matches_header(lib):
    from decentralized_ekf_mhe_b200.params import DekfConfig
    cfg = DekfConfig()
    assert lib.dekf_config_default_go1(C.byref(cfg)) == 0
    # parameters_go1.yaml values survive the round trip through the C struct layout
    assert cfg.abi_version == 3 and cfg.N == 20 and cfg.rate == 200 and cfg.num_legs == 4
    assert list(cfg.accel_bias_std) == [0.07, 0.02, 0.03]
    assert cfg.contact_effort_threshold == 150.0 and cfg.timeLimit == 0.0028
    assert list(cfg.ekf_quaternion_init) == [1.0, 0.0, 0.0, 0.0] and cfg.ekf_rate == 500
    assert lib.dekf_config_default_cassie(C.byref(cfg)) == 0 and cfg.num_legs == 2 and cfg.robot == 1
    assert lib.dekf_config_default_pogox(C.byref(cfg)) == 0 and cfg.num_legs == 1 and cfg.robot == 2


def test_create_rejects_bad_config(lib):
    from decentralized_ekf_mhe_b200.params import DekfConfig
    cfg = DekfConfig()
    lib.dekf_config_default_go1(C.byref(cfg))
    h = C.c_void_p()
    cfg.n_instances = 0
    assert lib.dekf_create(C.byref(cfg), C.byref(h)) == -1 and not h.value
    cfg.n_instances = 4
    cfg.num_legs = 3
    assert lib.dekf_create(C.byref(cfg), C.byref(h)) == -1
    cfg.num_legs = 4
    cfg.abi_version = 99
    assert lib.dekf_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.dekf_create(None, C.byref(h)) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from decentralized_ekf_mhe_b200.params import DekfConfig
    cfg = DekfConfig()
    lib.dekf_config_default_go1(C.byref(cfg))
    cfg.n_instances = 8
    h = C.c_void_p()
    assert lib.dekf_create(C.byref(cfg), C.byref(h)) == -2  # DEKF_ENODEV
    from decentralized_ekf_mhe_b200 import estimator
    with pytest.raises(estimator.DekfError):
        estimator.BatchedEstimator(estimator.robot_params("go1"), 8)


def test_product_package_does_not_import_oracle():
    """No import / include / dlopen of oracle/ or of the host debug harness anywhere in the product package."""
    pkg = os.path.join(ROOT, "decentralized_ekf_mhe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                continue
            for line in open(os.path.join(dirpath, f)):
                code = line.split("//")[0].split("#", 1)[0] if f.endswith(".py") else line.split("//")[0]
                if re.search(r"\b(import|from|include|CDLL|dlopen)\b", line):
                    assert not re.search(r"oracle|hostsim", code), (f, line)
