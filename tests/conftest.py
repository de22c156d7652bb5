"""pytest configuration: import paths, the `gpu` marker, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_ROOT = os.environ.get("KNPEMI_REFERENCE", "/root/reference")
MODEL_NAMES = ("hh_ideal", "hh_tissue", "glial_tissue", "glial_bench", "calibration", "hh_test")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Everything __graft_entry__.build() produces (idempotent, content-addressed)."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def reference_root():
    if not os.path.isdir(REFERENCE_ROOT):
        pytest.skip("reference tree not mounted (only present in the build container)")
    shim = os.path.join(ROOT, "tests", "shims")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    return REFERENCE_ROOT


def load_reference_module(reference_root, name):
    import importlib.util
    from knpemi_b200.models import REFERENCE_FILE
    path = os.path.join(reference_root, REFERENCE_FILE[name])
    spec = importlib.util.spec_from_file_location("reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
