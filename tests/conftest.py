import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


GOLDEN = ["short_s11", "short_s12_deep", "short_s13_jitter"]


def load_golden(name):
    with np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=GOLDEN)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope="session")
def engine():
    from stringtie_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()
