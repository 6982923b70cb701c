import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def oracle(built):
    import pna_oracle
    pna_oracle.lib()
    return pna_oracle


@pytest.fixture(scope="session")
def pna(built):
    return importlib.import_module("portable-network-archive_b200")


@pytest.fixture(scope="session")
def ctx(pna):
    c = pna.Context(0)   # raises when no sm_100 device is usable: GPU tests must not pass on a fallback
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    d = os.path.join(ROOT, "tests", "golden")
    m = json.load(open(os.path.join(d, "manifest.json")))
    m["dir"] = d
    return m
