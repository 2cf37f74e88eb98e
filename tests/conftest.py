import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=[pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def op_device(request, monkeypatch):
    """Device on which the mal_b200 operators run.  `emu` points the operator layer at the
    host-emulated twin of the kernels (tests/emu) so the autograd / dict plumbing is covered on
    CPU; this hook exists only here - the package itself refuses CPU tensors."""
    import torch
    if request.param == "emu":
        from mal_b200 import ops
        from tests.emu.emu_lib import emu
        handle = emu()
        monkeypatch.setattr(ops, "_lib", lambda t: handle)
        return torch.device("cpu")
    return torch.device("cuda:0")
