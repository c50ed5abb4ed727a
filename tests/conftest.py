import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = '/root/reference'


def reference_path():
    """Directory holding the UNMODIFIED reference orchestration layer: /root/reference in the build
    container, else the byte-identical copy staged by oracle/build_ref.py under oracle/_ref/ (git-ignored;
    it travels to the GPU box with the snapshot).  None when neither is there."""
    from oracle import build_ref
    return build_ref.path()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the unmodified reference L3 (/root/reference or oracle/_ref)")


def pytest_collection_modifyitems(config, items):
    have_ref = reference_path() is not None
    skip_ref = pytest.mark.skip(reason="neither /root/reference nor oracle/_ref (python oracle/build_ref.py) is present")
    for item in items:
        if 'reference' in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture
def reference_edrgp():
    """The UNMODIFIED reference orchestration layer (edrgp.edr / base / utils / datasets)."""
    ref = reference_path()
    if ref not in sys.path:
        sys.path.append(ref)
    import edrgp  # noqa: F401
    return edrgp
