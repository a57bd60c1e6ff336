import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure libvcpenc.so and the oracle exist (the driver normally ran build() already)."""
    from video_codec_pipeline_b200 import api
    from oracle import pyoracle
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    pyoracle.build()
    return True


CASES = [
    # w, h, frames, gop, slices, deblock_idc, qp
    (64, 48, 3, 60, 1, 1, 26),
    (64, 48, 3, 60, 1, 0, 26),
    (320, 180, 8, 4, 3, 0, 30),
    (320, 180, 8, 4, 3, 2, 18),
    (176, 144, 10, 5, 2, 0, 12),
    (176, 144, 10, 5, 9, 0, 51),
    (640, 360, 6, 60, 1, 0, 40),
    (208, 114, 5, 3, 1, 0, 33),   # ragged: neither dimension a multiple of 16
]
