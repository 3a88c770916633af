import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["sasrec_c1_small", "sasrec_l20_d64", "sasrec_dh128"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    g["grads"] = {k[len("grad/"):]: v for k, v in g.items() if k.startswith("grad/")}
    g["cfg"] = {k[len("cfg_"):]: int(v) for k, v in g.items() if k.startswith("cfg_")}
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
