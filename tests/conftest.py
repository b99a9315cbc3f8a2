import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def record_measured(name, value):
    """Appends a measured parity figure to gpurun_out/measured.jsonl (read back to set tolerances ~1.2x measured)."""
    import json

    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "measured.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, "value": value}) + "\n")
    except Exception:
        pass
    print(f"[measured] {name} = {value}")
