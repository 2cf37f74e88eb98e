#!/usr/bin/env python
"""Instruction / sample shares of a kernel between its barriers, from an ncu report captured with
--import-source on: SASS rows sorted by address, split at every BAR.SYNC / SYNCS wait.

    python tools/ncu_phases.py <report.ncu-rep> <kernel regex> [launch skip]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, kregex = sys.argv[1:3]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kregex, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    sass = {}
    for r in rows:
        if r and r[0] == "Address":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or len(r) < 8 or not r[0].startswith("0x"):
            continue
        try:
            sass[int(r[0], 16)] = (r[hdr["Source"]].strip(), float(r[hdr["Instructions Executed"]] or 0), float(r[hdr["# Samples"]] or 0))
        except (ValueError, KeyError):
            pass
    if not sass:
        print(raw[:2000])
        return
    ti = sum(v[1] for v in sass.values())
    ts = sum(v[2] for v in sass.values())
    print("total warp instructions %.4e, samples %d, SASS lines %d" % (ti, ts, len(sass)))
    seg_i = seg_s = 0.0
    n = 0
    ops = Counter()
    start = None
    for addr in sorted(sass):
        text, ni, ns = sass[addr]
        if start is None:
            start = addr
        seg_i += ni
        seg_s += ns
        n += 1
        op = text.split()[1] if text.startswith("@") else text.split()[0]
        ops[op.split(".")[0]] += ni
        if "BAR.SYNC" in text or "EXIT" in text or ("SYNCS" in text and "TRYWAIT" in text):
            if seg_i > 0.002 * ti or seg_s > 0.002 * ts:
                top = ", ".join("%s %.0f%%" % (k, 100 * v / max(seg_i, 1)) for k, v in ops.most_common(6))
                print("%#x..%#x %5d sass  inst %5.1f%%  samples %5.1f%%  ends at %-28s | %s" % (
                    start - min(sass), addr - min(sass), n, 100 * seg_i / ti, 100 * seg_s / max(ts, 1), text[:28], top))
            seg_i = seg_s = 0.0
            n = 0
            ops = Counter()
            start = None


if __name__ == "__main__":
    main()
