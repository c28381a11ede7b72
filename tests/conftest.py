import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    return load


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The .so is git-ignored: (re)build it in-tree when sources are newer (nvcc cross-compiles
    sm_100a without a GPU).  On the GPU box the prebuilt .so travels with the snapshot."""
    import shutil
    from vognet_pytorch_b200 import _lib
    if _lib.needs_build() and shutil.which('nvcc'):
        _lib.build()
    yield
