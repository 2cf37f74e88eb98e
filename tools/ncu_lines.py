#!/usr/bin/env python
"""Aggregate an ncu source-page capture per CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled function substring> [launch index]

ncu's `--page source --csv` lists SASS instructions with executed counts and stall samples but no
line numbers; `nvdisasm -g` lists the same SASS in the same order with `//## File ... line N`
markers.  Joining the two by instruction order gives instructions and stall samples per line.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_lines(cubin, func):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    out, cur_line, active = [], None, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = func in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur_line, m.group(2).strip()))
    return out


def main():
    rep, kregex, cubin, func = sys.argv[1:5]
    skip = sys.argv[5] if len(sys.argv) > 5 else "0"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    sass = sass_lines(cubin, func)
    if len(data) >= 2 * len(sass):   # ncu repeats the table (one per view); keep the first
        data = data[:len(sass)]
    if len(sass) != len(data):
        print("warning: %d SASS rows from ncu vs %d from nvdisasm" % (len(data), len(sass)), file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    tot_i = tot_s = 0.0
    for (line, _op), r in zip(sass, data):
        n = float(r[ix["Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        sh = float(r[ix.get("L1 Wavefronts Shared", 0)] or 0) if "L1 Wavefronts Shared" in ix else 0
        ex = float(r[ix["L1 Wavefronts Shared Excessive"]] or 0) if "L1 Wavefronts Shared Excessive" in ix else 0
        a = agg[line]
        a[0] += n; a[1] += s; a[2] += sh; a[3] += ex
        tot_i += n; tot_s += s
    print("total warp instructions %.3e, samples %d" % (tot_i, tot_s))
    print("%7s %7s %10s %10s  line" % ("inst%", "samp%", "smem_wf", "smem_exc"))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print("%6.2f%% %6.2f%% %10.0f %10.0f  %s:%s" % (100 * a[0] / tot_i, 100 * a[1] / max(tot_s, 1), a[2], a[3],
                                                       line[0] if line else "?", line[1] if line else "?"))


if __name__ == "__main__":
    main()
