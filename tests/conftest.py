import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_small(golden_dir):
    """{name: (input, reference output)} from the reference's own CPU path (make_golden.py)."""
    import numpy as np
    z = np.load(os.path.join(golden_dir, "lab_small.npz"))
    names = sorted(k[4:] for k in z.files if k.startswith("in__"))
    return {n: (z["in__" + n], z["out__" + n]) for n in names}


@pytest.fixture(scope="session")
def golden_mixed(golden_dir):
    import numpy as np
    z = np.load(os.path.join(golden_dir, "mixed_sign.npz"))
    names = sorted(k[4:] for k in z.files if k.startswith("in__"))
    return {n: (z["in__" + n], z["out__" + n]) for n in names}


@pytest.fixture(scope="session")
def golden_large(golden_dir):
    import json
    with open(os.path.join(golden_dir, "large.json")) as f:
        return json.load(f)
