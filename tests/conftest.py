import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A persistent (cooperative / cluster) kernel that deadlocks would hang the whole run: every GPU test
    gets a wall-clock limit when pytest-timeout is installed (the slowest one takes ~6 s on a B200)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(300, method="thread"))   # a hung CUDA call never returns to Python


def golden_cases():
    """Fixtures of the hot path (process_hessian_alt + gptq_fwrd); the chol_* / sketch_* fixtures of the
    other front ends have their own tests (test_oracle_frontends.py, test_gpu_frontends.py)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("chol_", "sketch_")))


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
        return cache[name]

    return load
