import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# The reference links zstd 1.5.7 (Cargo.lock: zstd-sys 2.0.14+zstd.1.5.7); the system libzstd of this image is 1.5.5, the
# same image also carries 1.5.7 inside pillow.libs.  The oracle uses 1.5.7 when it is there (accept / reject behaviour on
# malformed streams differs between the two: 1.5.6 started to reject non-zero reserved bits of the sequence modes byte).
for _d in [p for p in sys.path if p.endswith("site-packages")] + [os.path.join(sys.prefix, "lib", "python%d.%d" % sys.version_info[:2], "site-packages")]:
    _c = [f for f in (os.listdir(os.path.join(_d, "pillow.libs")) if os.path.isdir(os.path.join(_d, "pillow.libs")) else []) if f.startswith("libzstd") and "1.5.7" in f]
    if _c:
        os.environ.setdefault("PNA_ORACLE_LIBZSTD", os.path.join(_d, "pillow.libs", _c[0]))
        break


def pytest_addoption(parser):
    # development aid for a box without a GPU: run the kernels' logic under the fiber-based SIMT emulator of
    # tests/emu (test infrastructure; the package never loads it).  Parity evidence comes from `-m gpu` on a B200 only.
    parser.addoption("--emu", action="store_true", default=False, help="run -m gpu tests against tests/emu/_gen (SIMT emulator, debugging only)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow_emu: too large for the SIMT emulator (skipped under --emu)")
    if config.getoption("--emu"):
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        gen = build_emu.build()
        ffi = importlib.import_module("portable-network-archive_b200._ffi")
        host = importlib.import_module("portable-network-archive_b200._host")
        ffi.LIB_PATH = os.path.join(gen, "libpna_cuda.so")
        host.LIB_PATH = os.path.join(gen, "libpna_host.so")


def pytest_collection_modifyitems(config, items):
    if config.getoption("--emu"):
        skip = pytest.mark.skip(reason="too large for the SIMT emulator")
        for it in items:
            if "slow_emu" in it.keywords:
                it.add_marker(skip)


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
