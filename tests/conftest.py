import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the .so is git-ignored: build it in-tree when missing or stale (nvcc
    # cross-compiles sm_100a without a GPU)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_oc_b200_build", os.path.join(ROOT, "orthogonal_classifiers_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # by path: importing the package needs the library
    if mod.needs_build():
        mod.build_library()


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


def golden_sites(g, prefix):
    sites, i = [], 0
    while f"{prefix}_site_{i}" in g:
        sites.append(g[f"{prefix}_site_{i}"])
        i += 1
    return sites


@pytest.fixture(scope="session")
def gpu_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import orthogonal_classifiers_b200 as oc
    return oc
