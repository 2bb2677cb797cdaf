import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_ops(oracle):
    from tests.helpers import OracleOps
    return OracleOps(oracle)


@pytest.fixture(scope="session")
def gpu():
    """The product library on cuda:0.  Fails loudly (no CPU fallback) when the extension or the device is missing."""
    import latticefold_b200 as lf
    return lf
