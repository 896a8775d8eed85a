import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import orc

    orc.make()
    return orc.Port()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind oracle/_ref (prebuilt in the authoring container)."""
    from oracle import orc

    if not orc.Reference.available(True):
        pytest.skip("oracle/_ref not built (no /root/reference in this environment)")
    return orc.Reference(True)
