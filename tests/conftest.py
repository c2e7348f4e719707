import json
import os
import struct
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# ad-hoc differential campaigns on the GPU box: VMB_SEED_OFFSET=k shifts every seeded generator of the GPU tests
# (the committed expectations are for offset 0; a few structural assertions, e.g. "at least 50 index blocks", may not hold elsewhere)
SEED0 = int(os.environ.get("VMB_SEED_OFFSET", "0"))

STALE_NAN = struct.unpack("<d", struct.pack("<Q", 0x7FF0000000000002))[0]


def gofloat(s):
    """decode the float encoding used by tests/golden/go_kats.json"""
    if s == "nan":
        return float("nan")
    if s == "inf":
        return float("inf")
    if s == "-inf":
        return float("-inf")
    if s == "stale":
        return STALE_NAN
    return float.fromhex(s)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "go_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib
