import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/liboracle.so), built on demand."""
    from oracle import oracle_py

    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def strings_txt():
    return os.path.join(DATA, "strings.txt")


def golden_runs():
    """[(pcap name, proto, expected stdout without the Elapsed line)] from tests/golden/expected."""
    runs = []
    exp_dir = os.path.join(GOLDEN, "expected")
    for name in sorted(os.listdir(exp_dir)):
        pcap, proto, _ = name.rsplit(".", 2)
        with open(os.path.join(exp_dir, name), "rb") as f:
            runs.append((pcap, proto, f.read()))
    return runs
