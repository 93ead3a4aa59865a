import os
import sys

import pytest

# Launches with less than one face tile per resident warp take the plain tile kernel instead of the persistent bulk-copy
# pipeline (pdes_api.cu, launch_faces).  Nearly every parity case here is that small: keep the tests on the kernel the
# benchmark sizes run (test_face_kernel_variants_bitwise_equal covers the small-launch choice).
os.environ.setdefault("PDES_FACE_SMALL", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
