import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the reference build)")


def pytest_collection_modifyitems(config, items):
    from _libs import have_ref

    if have_ref():
        return
    skip = pytest.mark.skip(reason="oracle/_ref/libmcref.so not built (no /root/reference here)")
    for item in items:
        if "ref" in item.keywords:
            item.add_marker(skip)
