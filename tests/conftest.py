import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "omnirevolve-image-processor_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_DIR = "/root/reference/image_processor"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only; skipped elsewhere)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE_DIR)
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present on this box"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
