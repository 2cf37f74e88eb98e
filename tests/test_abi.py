"""The C-ABI library loads on a GPU-less host and exports every function include/mal_b200.h
declares; the ctypes table in mal_b200/_capi.py covers the same set; the Python structs have the
size the C compiler gives them.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from mal_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mal_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mal_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/mal_b200.h but not exported"
    assert set(names) == set(_capi.EXPORTS), set(names) ^ set(_capi.EXPORTS)
    _capi.bind(lib)
    assert lib.mal_abi_version() == _capi.ABI_VERSION


def test_struct_sizes_match_the_c_compiler(tmp_path):
    pairs = [("mal_photo_args", _capi.PhotoArgs), ("mal_cost_volume_args", _capi.CostVolumeArgs),
             ("mal_smooth_args", _capi.SmoothArgs), ("mal_main_terms_args", _capi.MainTermsArgs),
             ("mal_matching_mask_args", _capi.MatchingMaskArgs), ("mal_step_combine_args", _capi.StepCombineArgs),
             ("mal_forward_warp_args", _capi.ForwardWarpArgs), ("mal_corr_args", _capi.CorrArgs),
             ("mal_dynamic_instance_args", _capi.DynamicInstanceArgs), ("mal_temporal_args", _capi.TemporalArgs)]
    body = "".join('printf("%%zu\\n", sizeof(%s));' % name for name, _ in pairs)
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mal_b200.h"\nint main(void){' + body + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(s) for _, s in pairs]


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libmal_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.lib()


def test_bad_arguments_come_back_as_error_codes():
    """No exception or crash crosses the C boundary: a bad call returns a status and leaves a message
    (argument validation happens before any launch, so the host-emulated twin exercises the same code)."""
    import torch
    from mal_b200 import raw
    from tests.emu.emu_lib import emu
    h = emu()
    f = torch.rand(1, 6, 8, 8)            # 6 channels: not a multiple of 4
    with pytest.raises(RuntimeError, match="channel quads"):
        raw.corr_pyramid(h, f, 2)
    with pytest.raises(RuntimeError, match="levels do not fit"):
        raw.corr_pyramid(h, torch.rand(1, 8, 4, 4), 4)
    pyr = raw.corr_pyramid(h, torch.rand(1, 8, 8, 8), 2)
    with pytest.raises(RuntimeError, match="heads"):
        raw.corr_lookup(h, torch.rand(1, 8, 8, 8), pyr, torch.zeros(1, 2, 2, 3, 8, 8), num_head=3)
    with pytest.raises(ValueError, match="coords must be"):
        raw.corr_lookup(h, torch.rand(1, 8, 8, 8), pyr, torch.zeros(1, 2, 2, 3, 8, 9))
    tgt = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="WARP mode needs"):
        raw.photo(h, target=tgt, src=[tgt, tgt])                       # no depth / K / T
    with pytest.raises(RuntimeError, match="bad shape"):
        raw.photo(h, target=torch.rand(1, 3, 2, 2), src=[torch.rand(1, 3, 2, 2)] * 2, mode=raw.PHOTO_PRED)
    a = _capi.PhotoArgs()
    assert h.mal_photo_finalize(ctypes.byref(a), None) != 0 and b"mal_photo_finalize" in h.mal_last_error()
    assert h.mal_upsample_bilinear(None, 1, 4, 4, 8, 8, None, None) != 0
