#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep captured with
--import-source on:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; ncu_lines.py x.csv src.cuh"""
import csv
import sys


def num(v):
    try:
        return int(v)
    except (ValueError, TypeError):
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    src = open(sys.argv[2]).read().split("\n") if len(sys.argv) > 2 else None
    name = sys.argv[2].split("/")[-1] if src else None
    cur = hdr = None
    agg = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        agg.append((cur, ln, dict(zip(hdr, r))))
    tot = sum(num(d["Instructions Executed"]) for _, _, d in agg)
    tots = sum(num(d["# Samples"]) for _, _, d in agg)
    print("warp instructions %d, stall samples %d" % (tot, tots))
    for key in ("stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_mio", "stall_math", "stall_branch_resolving"):
        print("  %-24s %5.1f %%" % (key, 100.0 * sum(num(d[key]) for _, _, d in agg) / max(tots, 1)))
    print("top lines by samples:")
    for c, l, d in sorted(agg, key=lambda t: -num(t[2]["# Samples"]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
        text = src[l - 1].strip()[:72] if (src and c == name) else d["Source"][:72]
        print("  %-14s %4d  samples %5.1f%%  inst %5.1f%%  long_sb %5d barrier %5d wait %5d | %s" % (
            c, l, 100.0 * num(d["# Samples"]) / max(tots, 1), 100.0 * num(d["Instructions Executed"]) / max(tot, 1),
            num(d["stall_long_sb"]), num(d["stall_barrier"]), num(d["stall_wait"]), text))


if __name__ == "__main__":
    main()
