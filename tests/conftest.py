import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds libscht_b200.so on first use if it is missing (nvcc cross-compiles without a GPU)."""
    m = importlib.import_module("semi-convex_hull_tree_b200")
    if not os.path.exists(m.LIB_PATH):
        m.build_library()
    return m


@pytest.fixture(scope="session")
def orc():
    """CPU oracle (test infrastructure only)."""
    import oracle as o
    o.lib()
    return o


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["K"] = int(g["K"])
    g["pct"] = float(g["pct"])
    g["seed"] = int(g["seed"])
    g["tree"] = {k[5:]: (v if v.ndim else v.item()) for k, v in g.items() if k.startswith("tree_")}
    return g


@pytest.fixture(scope="session", params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    ne = bits(a) != bits(b)
    assert not ne.any(), "%s: %d of %d elements differ (first at %s: %r vs %r)" % (
        what, int(ne.sum()), ne.size, np.argwhere(ne)[0].tolist(), a[tuple(np.argwhere(ne)[0])], b[tuple(np.argwhere(ne)[0])])
