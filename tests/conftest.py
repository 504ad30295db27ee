import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the library consults its test / measurement switches (AGENDA_XRES, AGENDA_XSPLIT, ...) only when this was set before
# its first use; production runs (bench.py, the CLIs) never call getenv on the launch path
os.environ.setdefault("AGENDA_KNOBS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
