import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def import_reference():
    """Import the unmodified reference with the stand-ins for its missing third-party deps."""
    if not os.path.isdir(REFERENCE_DIR):
        pytest.skip("reference checkout not present (GPU box)")
    shims = os.path.join(ROOT, "oracle", "ref_shims")
    for p in (shims, REFERENCE_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gigagan_pytorch  # noqa: F401
    return gigagan_pytorch
