import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _tune_from_env():
    """GAB200_TEST_TUNE="3=8,6=1": run the whole suite with library tuning knobs (include/gab200_rasterizer.h
    GAB200_TUNE_*) set to non-default values -- how a candidate kernel variant is put through every parity test before
    it becomes the default."""
    spec = os.environ.get("GAB200_TEST_TUNE", "")
    if spec:
        from gaussianavatars_b200 import _native as N

        for kv in spec.split(","):
            k, v = kv.split("=")
            N.tune(int(k), int(v))
    yield
