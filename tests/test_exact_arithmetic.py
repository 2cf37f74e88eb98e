"""mal_common.cuh's xdivc<C> (three instructions instead of an IEEE division, used for the /9 of
avg_pool2d and the /3 of the channel means) is bit-identical to x / C.

The C twin of the sequence (tests/emu/exact_div_check.c, hardware fmaf) is checked against
x / C over every 61st float bit pattern by default (every exponent and sign, ~70 M values) and
over ALL 2^32 patterns with MAL_EXHAUSTIVE=1 (a few minutes; last run: 0 mismatches for every listed
constant, the only difference is the sign of a zero result for x = -0).  The same sequence divides by the
image-size constants of Project3D ((W-1), (H-1), W, H of the KITTI / Cityscapes shapes): those constants are
in the checked list as well, any other size keeps the IEEE division (mal_math.cuh size_div)."""
import os
import subprocess


def test_division_by_constant_is_exact(tmp_path):
    here = os.path.dirname(os.path.abspath(__file__))
    exe = tmp_path / "exact_div_check"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", os.path.join(here, "emu", "exact_div_check.c"),
                    "-o", str(exe), "-lm"], check=True)
    stride = "1" if os.environ.get("MAL_EXHAUSTIVE") == "1" else "61"
    res = subprocess.run([str(exe), stride], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.count("mismatches=0") == 12, res.stdout


def test_size_divisor_whitelist_matches_the_checked_constants():
    """size_div_verified (mal_math.cuh) may only list constants the exhaustive checker covers."""
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    cuh = open(os.path.join(os.path.dirname(here), "mal_b200", "csrc", "mal_math.cuh")).read()
    listed = {int(v) for v in re.search(r"static const int ok\[\] = \{([^}]*)\}", cuh).group(1).split(",")}
    chk = open(os.path.join(here, "emu", "exact_div_check.c")).read()
    checked = {int(float(v.strip().rstrip("f"))) for v in re.search(r"const float consts\[\] = \{([^}]*)\}", chk, re.S).group(1).split(",")}
    assert listed <= checked, listed - checked


import pytest  # noqa: E402


@pytest.mark.gpu
def test_packed_fp32_helpers_are_bit_exact_on_the_gpu(tmp_path):
    """FFMA2 / FADD2 / FMUL2 helpers == their scalar counterparts on sm_100a, including the ptxas
    mul+add contraction hazard documented in mal_common.cuh (tools/ubench/x2check.cu)."""
    import shutil
    here = os.path.dirname(os.path.abspath(__file__))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(os.path.dirname(here), "tools", "ubench", "x2check.cu")
    exe = str(tmp_path / "x2check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK"), res.stdout
