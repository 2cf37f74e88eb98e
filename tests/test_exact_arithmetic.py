"""mal_common.cuh's xdivc<C> (three instructions instead of an IEEE division, used for the /9 of
avg_pool2d and the /3 of the channel means) is bit-identical to x / C.

The C twin of the sequence (tests/emu/exact_div_check.c, hardware fmaf) is checked against
x / C over every 61st float bit pattern by default (every exponent and sign, ~70 M values) and
over ALL 2^32 patterns with MAL_EXHAUSTIVE=1 (about a minute; last run: 0 mismatches, the only
difference is the sign of a zero result for x = -0)."""
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
    assert res.stdout.count("mismatches=0") == 2, res.stdout
