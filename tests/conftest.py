"""pytest configuration: markers, import path, shared helpers.

`-m "not gpu"`: oracle vs the reference's golden vectors / committed fixtures, host-side logic, C-ABI
export checks, gloo world_size=2 sharding.  `-m gpu`: parity of the CUDA path (through the C-ABI)
against the oracle and the fixtures.
"""
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    return json.loads((GOLDEN / "reference_kats.json").read_text())


@pytest.fixture(scope="session")
def ref_fixtures():
    return json.loads((GOLDEN / "ref_fixtures.json").read_text())


@pytest.fixture(scope="session")
def api_sequences():
    return json.loads((GOLDEN / "ref_api_sequences.json").read_text())


@pytest.fixture(scope="session")
def harness():
    import oracle
    return oracle.Harness("port")


def gen_stream(harness, gen, k, n, literal=8):
    data = harness.generate(gen, k, 1, max(n, 1))[0].tobytes()[:n]
    if literal < 8:
        data = bytes(b & ((1 << literal) - 1) for b in data)
    return data
