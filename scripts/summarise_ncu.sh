#!/usr/bin/env bash
# Turn gpurun_out/prof_<what>.ncu-rep into the tracked summaries under profiles/ (raw metrics CSV + the
# hottest SASS lines).  Runs where the report is (no GPU needed).
set -uo pipefail
cd "$(dirname "$0")/.."
for what in ${@:-decode_linear gemm attn attnp gemm_fp8}; do
  src=$what; [ "$what" = decode_linear ] && src=decode
  rep=gpurun_out/prof_$src.ncu-rep
  [ -f "$rep" ] || { echo "missing $rep"; continue; }
  ncu -i "$rep" --page raw --csv > profiles/prof_$what.raw.csv 2>/dev/null
  ncu -i "$rep" --page source --csv 2>/dev/null | python scripts/summarise_ncu.py "$what"
  echo "profiles/prof_$what.raw.csv: $(wc -l < profiles/prof_$what.raw.csv) lines"
done
