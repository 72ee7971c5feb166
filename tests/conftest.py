"""pytest plumbing.  `-m "not gpu"`: oracle vs golden vectors, host logic, ABI
export check, gloo world_size-2 sharding.  `-m gpu`: the parity tests proper,
every one of them through the C ABI of libsclgpu.so."""
from __future__ import annotations

import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def port(oracle_mod):
    return oracle_mod.PortOracle()


@pytest.fixture(scope="session")
def ref(oracle_mod):
    if not oracle_mod.ref_available() and not os.path.isdir("/root/reference/src/scl"):
        pytest.skip("oracle/_ref/libsclref.so not present")
    return oracle_mod.RefOracle()


@pytest.fixture(scope="session")
def orc(oracle_mod):
    """best available checker: the compiled reference if its .so travelled, else the port"""
    return oracle_mod.best_oracle()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(REPO, "tests", "golden", "scl_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def ctx(pkg):
    """One libsclgpu context on cuda:0.  Fails (never skips, never falls back)
    when the library or the GPU is missing."""
    c = pkg.Context(0)
    yield c
    c.close()
