import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from kb2e_oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from kb2e_oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libkb2e_ref.so not built (no /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "kb2e_golden.npz"))


def golden_case(golden, ci):
    model, D, dist = (int(x) for x in golden["cases"][ci])
    k = f"c{ci}_"
    g = {name[len(k):]: golden[name] for name in golden.files if name.startswith(k)}
    g["model"], g["D"], g["dist"] = model, D, dist
    g["wq"] = None if model == 0 else g["w"]
    return g


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU; fails (not skips) when it is missing on a GPU box."""
    import kb2e_b200
    return kb2e_b200.load_library()
