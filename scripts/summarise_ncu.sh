#!/usr/bin/env bash
# Turn gpurun_out/prof_<what>.ncu-rep into the tracked summaries under profiles/ (raw metrics CSV + the
# hottest SASS lines).  Runs where the report is (no GPU needed).
set -uo pipefail
cd "$(dirname "$0")/.."
for what in ${@:-decode_linear gemm attn}; do
  src=$what; [ "$what" = decode_linear ] && src=decode
  rep=gpurun_out/prof_$src.ncu-rep
  [ -f "$rep" ] || { echo "missing $rep"; continue; }
  ncu -i "$rep" --page raw --csv > profiles/prof_$what.raw.csv 2>/dev/null
  ncu -i "$rep" --page source --csv 2>/dev/null | python - "$what" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
what = sys.argv[1]
if len(rows) > 1:
    hdr = rows[0]
    try:
        si = next(i for i, h in enumerate(hdr) if h.strip().lower() in ("source", "sass"))
        ci = next(i for i, h in enumerate(hdr) if "Sampling" in h and "All" in h) if any("Sampling" in h for h in hdr) else None
    except StopIteration:
        si, ci = 0, None
    body = rows[1:]
    if ci is not None:
        def val(r):
            try:
                return float(r[ci])
            except Exception:
                return 0.0
        tot = sum(val(r) for r in body) or 1.0
        top = sorted(body, key=val, reverse=True)[:25]
        with open(f"profiles/prof_{what}.hot_sass.txt", "w") as f:
            f.write(f"# top SASS lines by warp-stall samples (all), total {tot:.0f}\n")
            for r in top:
                f.write(f"{100 * val(r) / tot:5.1f}%  {r[si]}\n")
PY
  echo "profiles/prof_$what.raw.csv: $(wc -l < profiles/prof_$what.raw.csv) lines"
done
