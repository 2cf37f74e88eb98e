#!/usr/bin/env python
"""Run the kernel tests on an AddressSanitizer build of the CPU twin (tests/emu).

compute-sanitizer is closed on the GPU pool, so out-of-bounds reads / writes of the kernels are
hunted here instead: the same .cu sources, compiled by g++ -fsanitize=address against the CUDA
emulator, executed by the normal parity tests (global buffers are torch CPU tensors, shared memory
is a heap vector, so ASan sees both).

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 python tools/asan_emu.py
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = "/tmp/mal_b200_emu_asan/libmal_b200_emu.so"


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-DMAL_EMU", "-ffp-contract=off", "-mfma",
           "-fno-fast-math", "-Wno-unused", "-fsanitize=address", "-fno-omit-frame-pointer", "-I",
           os.path.join(ROOT, "tests", "emu"), "-x", "c++"] + sorted(glob.glob(os.path.join(ROOT, "mal_b200", "csrc", "*.cu"))) + ["-o", OUT]
    subprocess.run(cmd, check=True)


def main():
    if "libasan" not in os.environ.get("LD_PRELOAD", ""):
        raise SystemExit(__doc__)
    build()
    import tests.emu.emu_lib as E
    from mal_b200 import _capi
    E._handle = _capi.bind(ctypes.CDLL(OUT))
    import pytest
    tests = ["test_cost_volume.py", "test_photo_kernel.py", "test_pointwise_kernels.py", "test_forward_warp.py",
             "test_dyn_utils.py", "test_step.py", "test_corr.py", "test_upsample.py", "test_temporal.py"]
    # (test_api.py is left out: any C++ exception thrown inside libtorch aborts under a preloaded ASan
    #  - an interception problem that has nothing to do with the kernels)
    sys.exit(pytest.main(["-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] +
                         [os.path.join(ROOT, "tests", t) for t in tests]))


if __name__ == "__main__":
    main()
