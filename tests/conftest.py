import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def psx_env(monkeypatch):
    """psx_env(NAME=value, OTHER=None) sets / clears libpsx's PSX_* kernel-selection switches and has the library read
    them again (it caches them at the first launch, include/psx.h: psx_reload_env); undone after the test."""
    from samplers_b200 import _native

    def apply(**kv):
        for k, v in kv.items():
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, str(v))
        _native.reload_env()

    yield apply
    monkeypatch.undo()
    _native.reload_env()
