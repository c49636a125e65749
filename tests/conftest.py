import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TESTDATA = os.path.join(ROOT, "testdata")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def testdata():
    return TESTDATA


@pytest.fixture(scope="session")
def oracle_models():
    """name -> OracleModel for the four shipped .matok files (CPU oracle, test infrastructure)."""
    from oracle import pyoracle
    return {n: pyoracle.OracleModel(os.path.join(TESTDATA, n))
            for n in ("tokenizer_de.matok", "tokenizer_en.matok", "simpletok.matok", "clitic_test.matok")}
