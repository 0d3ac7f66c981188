#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from a .ncu-rep
(needs -lineinfo and --import-source on). Usage: python tools/ncu_lines.py <rep> <kernel regex> [top]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines, tot_i, tot_s = [], 0, 0
    for r in rows:
        if len(r) >= 9 and r[0].isdigit() and r[2] == "-":
            inst, samp = int(r[7] or 0), int(r[6] or 0)
            lines.append((int(r[0]), r[1].strip()[:100], inst, samp))
            tot_i += inst
            tot_s += samp
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%5s %7s %7s  %s" % ("line", "inst%", "samp%", "source"))
    for ln, src, inst, samp in sorted(lines, key=lambda x: -x[2])[:top]:
        print("%5d %7.2f %7.2f  %s" % (ln, 100.0 * inst / max(tot_i, 1), 100.0 * samp / max(tot_s, 1), src))


if __name__ == "__main__":
    main()
