import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def hook_path(name):
    """Path of a reference hook file, or skip: /root/reference here, baseline/_ref/hooks on the GPU box."""
    from mpv_prescalers_b200.hookfile import HookError, find_hook

    try:
        return find_hook(name)
    except HookError:
        pytest.skip(f"reference hook {name} not available")


@pytest.fixture(scope="session")
def hooks():
    return hook_path
