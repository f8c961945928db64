import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: d[k] for k in d.files}
    for k in ("fit", "field"):
        if k in out:
            out[k] = str(out[k])
    for k in ("prior", "norm"):
        if k in out:
            out[k] = bool(out[k])
    return out


def fit_params_of(g):
    return {"initial_guess": [float(v) for v in g["x0"]],
            "param_bounds": [(float(a), float(b)) for a, b in g["bounds"]],
            "solver": "L-BFGS-B", "options": {}}


@pytest.fixture(scope="session")
def gpu_lib():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fetal_t2mapping_b200 as t2
    t2.init(0)
    return t2
