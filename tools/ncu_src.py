#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall shares from an ncu report captured with --import-source on.

    python tools/ncu_src.py <report.ncu-rep> <kernel regex> [launch skip] [top N]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kregex = sys.argv[1:3]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kregex, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr = None, None
    agg = defaultdict(lambda: defaultdict(float))
    text = {}
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r) if n not in ("Source",)}
            src_i = [i for i, n in enumerate(r) if n == "Source"]
            continue
        if hdr is None or len(r) < 10 or not r[0]:
            continue
        key = (fname, int(r[0]))
        text[key] = r[src_i[0]].strip()[:90]
        for n in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "stall_long_sb",
                  "stall_barrier", "stall_wait", "stall_short_sb", "stall_lg", "stall_mio", "stall_no_inst", "stall_math",
                  "L2 Theoretical Sectors Local"):
            if n in hdr and r[hdr[n]]:
                try:
                    agg[key][n] += float(r[hdr[n]])
                except ValueError:
                    pass
    ti = sum(a["Instructions Executed"] for a in agg.values())
    ts = sum(a["# Samples"] for a in agg.values())
    print("total warp instructions %.4e, samples %d" % (ti, ts))
    print("%6s %6s %9s %8s %6s %6s %6s %6s  line" % ("inst%", "samp%", "smem_wf", "excess", "longsb", "barr", "wait", "local"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        print("%5.2f%% %5.2f%% %9.0f %8.0f %6.0f %6.0f %6.0f %6.0f  %s:%d  %s" % (
            100 * a["Instructions Executed"] / max(ti, 1), 100 * a["# Samples"] / max(ts, 1), a["L1 Wavefronts Shared"],
            a["L1 Wavefronts Shared Excessive"], a["stall_long_sb"], a["stall_barrier"], a["stall_wait"],
            a["L2 Theoretical Sectors Local"], key[0], key[1], text[key]))


if __name__ == "__main__":
    main()
