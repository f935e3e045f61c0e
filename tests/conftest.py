import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (dev container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)


def pytest_sessionstart(session):
    """Make sure the in-tree C-ABI library matches the sources (no-op when the stamp is fresh; nvcc cross-compiles
    without a GPU).  A failed build is not hidden: the ops raise when the library is missing."""
    try:
        from uda_clr_b200 import build
        build.build()
    except Exception as e:  # pragma: no cover
        sys.stderr.write("conftest: could not (re)build libclr_b200.so: %s\n" % str(e)[:300])
