#!/usr/bin/env python
"""Aggregate a .ncu-rep's source page by CUDA source line: warp-level instructions executed and
stall samples per file:line (needs -lineinfo and --import-source on).

    python profiles/ncu_by_line.py gpurun_out/prof.ncu-rep [top_n] [kernel-name substring]
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, h, agg, text = None, None, {}, {}
    seen_kernel = None
    skipping = False
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            if want is not None and seen_kernel is None and want not in r[1]:
                h = None
                skipping = True
                continue
            skipping = False
            if seen_kernel is None:
                seen_kernel = r[1]
            elif r[1] != seen_kernel:
                break                       # one kernel only (the first, or the first whose name matches)
            continue
        if r[0] == "Line No":
            if skipping:
                continue
            h = r
            ie, si = h.index("Instructions Executed"), h.index("# Samples")
            continue
        if skipping or h is None or len(r) != len(h) or not r[0].isdigit():
            continue
        key = (path, int(r[0]))
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[ie] or 0)
        a[1] += int(r[si] or 0)
        text.setdefault(key, r[1].strip())
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("# %s\n# %s\n# warp-level instructions %d, stall samples %d" % (rep, (seen_kernel or "")[:100], ti, ts))
    print("# by instructions executed")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-18s:%-5d inst %10d %5.1f%%  samples %6d %5.1f%%  %s" % (key[0], key[1], a[0], 100.0 * a[0] / ti,
                                                                    a[1], 100.0 * a[1] / ts, text[key][:90]))
    print("# by file")
    byf = {}
    for key, a in agg.items():
        b = byf.setdefault(key[0], [0, 0])
        b[0] += a[0]
        b[1] += a[1]
    for f, b in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print("%-24s inst %10d %5.1f%%  samples %6d %5.1f%%" % (f, b[0], 100.0 * b[0] / ti, b[1], 100.0 * b[1] / ts))


if __name__ == "__main__":
    main()
