import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a CPU-only box they are skipped (instead of 80 errors that bury a real CPU
    failure).  On a GPU box nothing is skipped -- a missing libmonoloco_b200.so there must fail loudly."""
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the GPU box)")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
