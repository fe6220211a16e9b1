import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.build_oracle()
    return _orc.Oracle("orc")


@pytest.fixture(scope="session")
def ref():
    import orc as _orc
    _orc.build_oracle()
    if not _orc.have_ref():
        pytest.skip("oracle/_ref/libapus_ref.so absent (no /root/reference to build it from)")
    return _orc.Oracle("ref")
