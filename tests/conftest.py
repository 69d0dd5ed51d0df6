import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(autouse=True)
def _fresh_model_manager():
    """Every test starts with an empty model registry."""
    try:
        from elektronn2_b200.neuromancer import model_manager
        model_manager.reset()
    except Exception:
        pass
    yield
