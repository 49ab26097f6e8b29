import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = os.path.join(GOLDEN, "ref_data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def golden_counts():
    with open(os.path.join(GOLDEN, "counts_golden.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_cli():
    with open(os.path.join(GOLDEN, "cli_fixture.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_seq():
    with open(os.path.join(GOLDEN, "seq_golden.json")) as f:
        return json.load(f)
