import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from kmc_testlib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from kmc_testlib import Reference, ensure_reference_built
    if not ensure_reference_built():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Reference()
