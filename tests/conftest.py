import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(autouse=True)
def _chdir_root(monkeypatch):
    # the .cfg files name data/example_inlet.csv relative to the repo root,
    # exactly as the reference's configs do
    monkeypatch.chdir(ROOT)
