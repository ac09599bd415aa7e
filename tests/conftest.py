import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds (if needed) and loads libtzddpc.so; GPU tests must run the native path."""
    from tzddpc_b200 import build, _abi
    if not os.environ.get("TZ_SKIP_BUILD"):          # (experiments that ship a library built with extra flags)
        build.build(verbose=False)
    return _abi.lib()
