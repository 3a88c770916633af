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


def load_bench_shape_golden():
    """tests/golden/sasrec_bench_shape.npz (D=512, L=20, h=4, 2 layers -- the shape bench.py times): parameters are
    regenerated from the seed (oracle.make_golden.bench_shape_params), gradients are the stored subset (grad_subset)."""
    import numpy as np
    from oracle.make_golden import bench_shape_params
    z = np.load(os.path.join(GOLDEN_DIR, "sasrec_bench_shape.npz"))
    g = {k: z[k] for k in z.files}
    g["cfg"] = {k[len("cfg_"):]: int(v) for k, v in g.items() if k.startswith("cfg_")}
    shapes = {k[len("shape/"):]: tuple(int(x) for x in v) for k, v in g.items() if k.startswith("shape/")}
    g["params"] = bench_shape_params(shapes, g["cfg"]["seed"])
    g["grads"] = {k[len("grad/"):]: v for k, v in g.items() if k.startswith("grad/")}
    return g
