import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref built from /root/reference (dev container only)")


def pytest_collection_modifyitems(config, items):
    """Without a GPU the `gpu` tests are skipped (not failed): the product has no CPU path."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="needs a CUDA device (the library has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)
