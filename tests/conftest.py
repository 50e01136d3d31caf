import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def cv2_dispatch() -> str:
    """The lumina ``cv_dispatch`` value that names the mode the LIVE cv2 of this process is in right now
    (cv2.setUseOptimized is process-global and some fixtures switch it off): tests that compare with a live OpenCV
    call pass this to the product so that both sides run the same float path of adaptiveThreshold."""
    import cv2

    return "avx2" if cv2.useOptimized() and cv2.checkHardwareSupport(10) and cv2.checkHardwareSupport(12) else "plain"

