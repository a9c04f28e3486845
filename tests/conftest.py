import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def quiet_stderr():
    """context manager factory silencing fd 2 (the reference prints warnings there)"""
    import contextlib

    @contextlib.contextmanager
    def _quiet():
        sys.stderr.flush()
        saved = os.dup(2)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        try:
            yield
        finally:
            os.dup2(saved, 2)
            os.close(devnull)
            os.close(saved)

    return _quiet
