import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    # The product has no CPU fallback; even the CPU suite needs the shared
    # library to exist so that the ABI tests can load it (nvcc cross-compiles).
    lib = os.path.join(ROOT, "mhaq_b200", "csrc", "libmhaq_fq.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
