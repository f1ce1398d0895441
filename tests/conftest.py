import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def _gpu_available() -> bool:
    try:
        import dlv3p_b200
        return dlv3p_b200.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope='session')
def gpu():
    """GPU tests must FAIL (not skip) when selected with -m gpu and the native path is unavailable."""
    import dlv3p_b200
    dlv3p_b200.load_library()
    n = dlv3p_b200.device_count()
    assert n > 0, 'no CUDA device visible: -m gpu tests need a B200 (no CPU fallback exists)'
    info = dlv3p_b200.device_info(0)
    assert info['sm'][0] == 10, 'sm_100 device required, found sm_%d%d' % info['sm']
    return info
