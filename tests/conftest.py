import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


from brotli_libs import zeroed_malloc  # noqa: E402

zeroed_malloc()   # the reference is not deterministic on a dirty heap (see brotli_libs.zeroed_malloc)
