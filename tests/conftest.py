import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

BLOBS = ROOT / "hsr_env_b200" / "blobs"
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def models():
    from hsr_env_b200.model import Model

    return {p.stem: Model.load(p) for p in sorted(BLOBS.glob("*.hsrb"))}


@pytest.fixture(scope="session")
def ports(models):
    """fp64/fp32 C++ port of the substep (oracle side), one per model."""
    from oracle import port

    port.build()
    return {k: port.CpuPort(m) for k, m in models.items()}
