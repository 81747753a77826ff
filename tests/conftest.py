import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _build_once():
    import __graft_entry__ as entry
    entry.build()


@pytest.fixture(scope="session")
def built():
    _build_once()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference(built):
    """The unmodified reference build; only exists where oracle/_ref was compiled or travelled to."""
    from oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libdivquant_ref.so not present")
    return Reference()


@pytest.fixture(scope="session")
def pkg(built):
    return importlib.import_module("clusteringsegmentation-1_b200")


@pytest.fixture(scope="session")
def dq(pkg):
    """The CUDA path through the C ABI. Only for gpu-marked tests: the library aborts without a device."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a GPU"
    return pkg.DivQuant(timings=False)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "divquant_kat.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def images():
    return {name: np.load(os.path.join(GOLDEN, f"{name}_px.npz"))["px"].ravel() for name in ("batman", "cookie")}


def small_case_ids(golden_npz):
    return sorted({int(k[5:].split("_")[0]) for k in golden_npz.files if k.startswith("small")})
