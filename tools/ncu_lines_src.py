#!/usr/bin/env python
"""Stall samples of an .ncu-rep (captured with --import-source on) aggregated by CUDA source line.
Usage: python tools/ncu_lines_src.py rep.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
names = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: hdr.index(c) for c in names}
cur, agg = None, {}
for r in rows[h + 1:]:
    if r and r[0] != "":
        cur = (r[0], r[1][:86]); continue
    try:
        s = int(r[isamp] or 0); e = int(r[iex] or 0)
    except (ValueError, IndexError):
        continue
    a = agg.setdefault(cur, [0, 0, {}])
    a[0] += s; a[1] += e
    for c, i in idx.items():
        try: a[2][c] = a[2].get(c, 0) + int(r[i] or 0)
        except (ValueError, IndexError): pass
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    main = sorted(v[2].items(), key=lambda kv: -kv[1])[:2]
    print(k[0].rjust(4), str(v[0]).rjust(7), "%5.1f%%" % (100 * v[0] / tot), str(v[1]).rjust(10), k[1].ljust(86),
          " ".join("%s=%d" % (n[6:], c) for n, c in main))
