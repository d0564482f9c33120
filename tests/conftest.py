import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def gpu_ctx():
    from chord_b200 import runtime
    ctx = runtime.Context(0)        # raises loudly when the CUDA library or the device is missing
    yield ctx
    ctx.close()
