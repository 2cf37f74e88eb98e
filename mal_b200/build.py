"""Build libmal_b200.so (sm_100a) in-tree with nvcc.  `python -m mal_b200.build`."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmal_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "mal_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source (one nvcc per file, in parallel) and link them into one shared library
    next to the package."""
    if not force and not needs_build():
        return LIB
    import concurrent.futures
    import tempfile
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f not in ("--use_fast_math=false", "-shared")]
    objdir = tempfile.mkdtemp(prefix="mal_b200_obj_")

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        res = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, res

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    log = "".join(r.stderr for _, r in results)
    failed = [r for _, r in results if r.returncode != 0]
    if verbose or failed:
        sys.stderr.write("".join(r.stdout + r.stderr for _, r in (results if verbose else [(None, f) for f in failed])))
    if failed:
        raise RuntimeError("nvcc failed building libmal_b200.so")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] +
                         [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libmal_b200.so")
    with open(os.path.join(PKG, "csrc", "ptxas.log"), "w") as f:
        f.write(log)
    import shutil
    shutil.rmtree(objdir, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
