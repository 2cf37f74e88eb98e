"""Parametrisation helper: run the same parity test through the host-emulated twin
(CPU suite) and through the real CUDA library (`-m gpu`)."""
import pytest
import torch

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def handle_and_device(backend):
    if backend == "emu":
        from tests.emu.emu_lib import emu
        return emu(), torch.device("cpu")
    from mal_b200 import _capi
    return _capi.lib(), torch.device("cuda:0")
