import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def workloads():
    import kxpu_b200  # noqa: F401
    from kxpu_b200 import workloads as W
    return W


@pytest.fixture(scope="session")
def pci_text(workloads):
    t = workloads.load_pci_ids()
    assert hashlib.sha256(t).hexdigest() == workloads.PCI_IDS_SHA256
    return t


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="session")
def oracle_rows(oracle, pci_text):
    return oracle.table_build(pci_text)


@pytest.fixture(scope="session")
def kx():
    """One GPU context for the whole session.  Fails (does not skip) when the CUDA library
    or the GPU is missing: the product has no CPU path."""
    import kxpu_b200 as K
    k = K.Kxpu(0)
    yield k
    k.close()
