import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    from tests.golden.cases import build_cases
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    cases = {c["name"]: c for c in build_cases()}
    return man, cases


def golden_rfq(name):
    return open(os.path.join(ROOT, "tests", "golden", name + ".rfq"), "rb").read()
