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
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mal_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(mal_photo_args),sizeof(mal_cost_volume_args),sizeof(mal_smooth_args),"
                   "sizeof(mal_main_terms_args),sizeof(mal_matching_mask_args),sizeof(mal_step_combine_args),"
                   "sizeof(mal_forward_warp_args));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(s) for s in (_capi.PhotoArgs, _capi.CostVolumeArgs, _capi.SmoothArgs,
                                       _capi.MainTermsArgs, _capi.MatchingMaskArgs, _capi.StepCombineArgs,
                                       _capi.ForwardWarpArgs)]
    assert got == want


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libmal_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.lib()
