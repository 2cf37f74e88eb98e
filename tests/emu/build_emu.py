"""Build the host-emulated twin of libmal_b200 for the CPU-only test-suite.

TEST INFRASTRUCTURE.  The same mal_b200/csrc/*.cu files are compiled by g++ with -DMAL_EMU,
which swaps <cuda_runtime.h> for tests/emu/cuda_emu.h (fibers for CUDA threads).  The result
(tests/emu/_build/libmal_b200_emu.so) exports the C ABI of include/mal_b200.h but takes host
pointers.  It is loaded only by tests (tests/emu/emu_lib.py); the mal_b200 package never
looks for it - there is no CPU fallback in the product.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mal_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libmal_b200_emu.so")


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cuda_emu.h"),
                                                             os.path.join(ROOT, "include", "mal_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DMAL_EMU", "-ffp-contract=off", "-mfma",
           "-fno-fast-math", "-Wno-unused", "-I", HERE, "-x", "c++"] + srcs + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulator twin")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
