"""Shared test plumbing: the ``gpu`` marker and import paths.

``-m "not gpu"`` tests run on a CPU-only box: they hold the oracle to the golden
vectors made from the reference's own source, exercise host-side logic and check
that the C-ABI library loads and exports every declared symbol.  ``-m gpu`` tests
are the parity tests proper: they call the CUDA path through the C-ABI and compare
it with the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """GPU tests must FAIL, not skip, on a GPU box without a working library; on a box
    without a device they are skipped only when the user did not ask for them with -m gpu."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device on this box')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return load
