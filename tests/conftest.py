import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def sp():
    """The product package (ctypes binding over libstark_b200.so)."""
    import build_ext
    build_ext.build()
    return importlib.import_module("stark-prover_b200")


@pytest.fixture(scope="session")
def ctx(sp):
    """A context on cuda:0 for the default field; GPU tests only."""
    c = sp.Context(sp.P_DEFAULT, sp.G_DEFAULT, 0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    out = {}
    gdir = os.path.join(ROOT, "tests", "golden")
    for f in os.listdir(gdir):
        if f.endswith(".json"):
            out[f[:-5]] = json.load(open(os.path.join(gdir, f)))
    return out
