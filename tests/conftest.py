import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box")
    config.addinivalue_line("markers", "slow: multi-minute CPU test (24-layer oracles); still part of the default run")


@pytest.fixture(scope="session")
def s3b_lib():
    """The C-ABI shared library; built on demand when nvcc is available."""
    from s3prl_b200 import build, lib

    if not build.is_current():
        build.build()
    return lib.load()
