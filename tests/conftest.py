import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle (test infrastructure)."""
    import oracle_binding
    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def rt():
    """The product package; builds libraytracer.so in-tree if it is missing or stale."""
    build = importlib.import_module("rust-swift-raytracer_b200.build")
    build.build()
    return importlib.import_module("rust-swift-raytracer_b200")


@pytest.fixture(scope="session")
def scenes():
    return importlib.import_module("rust-swift-raytracer_b200.scenes")


@pytest.fixture(scope="session")
def gpu_rt(rt):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box "
                    "(the library has no CPU render path to fall back to)")
    return rt
