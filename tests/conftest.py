import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (absent on the GPU box)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(os.path.join(REFERENCE, 'edrgp'))
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if 'reference' in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture
def reference_edrgp():
    """The UNMODIFIED reference orchestration layer (edrgp.edr / base / utils / datasets)."""
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)
    import edrgp  # noqa: F401
    return edrgp
