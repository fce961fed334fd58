#!/usr/bin/env bash
# Race / memory checking of the hand-written kernels (the reference has no sanitizer story;
# its only concurrency is Python threads).  Runs the single-GPU kernel tests under compute-sanitizer:
#   memcheck  - out-of-bounds / misaligned accesses (incl. the bulk-copy rings and TMA tiles)
#   racecheck - shared-memory hazards between the warps of a CTA (mbarrier rings, split-KV merge)
#   synccheck - illegal barrier use (divergent __syncthreads, mbarrier misuse)
# The cross-GPU flag protocol is outside what the tools model; it is covered by the multi-GPU tests.
# usage: scripts/sanitize.sh [memcheck|racecheck|synccheck|all] [pytest -k expression]
set -uo pipefail
cd "$(dirname "$0")/.."
tool=${1:-memcheck}
expr=${2:-"linear or attn or embed or sample or gemm"}
out=${SANITIZE_OUT:-gpurun_out}
mkdir -p "$out"
tools=("$tool")
[ "$tool" = all ] && tools=(memcheck racecheck synccheck)
rc=0
for t in "${tools[@]}"; do
  log="$out/sanitize_$t.log"
  echo "== compute-sanitizer --tool $t  (-k \"$expr\") -> $log"
  timeout 1500 compute-sanitizer --tool "$t" --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -x -q -k "$expr" -p no:cacheprovider > "$log" 2>&1
  r=$?
  tail -4 "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -2
  [ $r -ne 0 ] && rc=$r
done
exit $rc
