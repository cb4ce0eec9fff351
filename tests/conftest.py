"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu and call the product only
through the C-ABI (libsmj.so); the oracle is the checker, never the thing under test."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_csv(tmp_path_factory):
    """Materialises a golden input (gz or plain) as a real file path; returns a getter."""
    d = tmp_path_factory.mktemp("golden_csv")

    def get(name):
        plain = os.path.join(GOLDEN, name)
        if os.path.exists(plain):
            return plain
        out = os.path.join(d, name)
        if not os.path.exists(out):
            with gzip.open(plain + ".gz", "rb") as g, open(out, "wb") as f:
                f.write(g.read())
        return out

    return get


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    oracle.build(ref=os.path.exists("/root/reference/sort-merge-join/cpu_app.c"))
    return oracle.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle
    if os.path.exists("/root/reference/sort-merge-join/cpu_app.c"):
        oracle.build(ref=True)
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference absent)")
    return oracle.Ref()
