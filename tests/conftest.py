import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "unverified: GPU engine written without GPU access and not yet run on a B200; "
                                       "skipped unless B2_RUN_UNVERIFIED=1 (first GPU call of the next session)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("B2_RUN_UNVERIFIED") == "1":
        return
    skip = pytest.mark.skip(reason="engine not yet verified on a B200 (set B2_RUN_UNVERIFIED=1 to run)")
    for item in items:
        if "unverified" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
