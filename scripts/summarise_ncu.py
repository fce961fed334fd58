#!/usr/bin/env python3
"""profiles/prof_<what>.hot_sass.txt from `ncu -i <rep> --page source --csv` on stdin."""
import csv
import sys

what = sys.argv[1]
rows = list(csv.reader(sys.stdin))
kernel = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
hi = next(i for i, r in enumerate(rows) if "Source" in r)
hdr, body = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
si = hdr.index("Source")
ci = next(i for i, h in enumerate(hdr) if h.startswith("Warp Stall Sampling (All"))


def val(r):
    try:
        return float(r[ci])
    except ValueError:
        return 0.0


tot = sum(val(r) for r in body) or 1.0
with open(f"profiles/prof_{what}.hot_sass.txt", "w") as f:
    f.write(f"kernel: {kernel}\ntotal warp-stall samples (first captured launch): {tot:.0f}\n")
    for r in sorted(body, key=val, reverse=True)[:20]:
        f.write(f"{val(r):6.0f} {100 * val(r) / tot:5.1f}%  {r[si].strip()}\n")
