import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) every native artefact once per session; see __graft_entry__.build()."""
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture()
def fixture_dir(tmp_path):
    """Scratch copy of the reference's test inputs + goldens (tests/golden/reference_fixtures.tar.gz)."""
    import runner
    return runner.unpack_fixtures(str(tmp_path))
