#!/usr/bin/env bash
# Race / memory checking of the hand-written kernels (the reference has no sanitizer story;
# its only concurrency is Python threads).  Runs the single-GPU kernel tests under compute-sanitizer:
#   memcheck  - out-of-bounds / misaligned accesses (incl. the bulk-copy rings and TMA tiles)
#   racecheck - shared-memory hazards between the warps of a CTA (mbarrier rings, split-KV merge)
#   synccheck - illegal barrier use (divergent __syncthreads, mbarrier misuse)
# The cross-GPU flag protocol itself is outside what the tools model (they see one device); `hop` runs the kernels
# that carry it — hop wait / signal / row copy / poison, tickets — under racecheck + synccheck on one GPU, and the
# multi-GPU tests cover the protocol end to end.
# usage: scripts/sanitize.sh [memcheck|racecheck|synccheck|all|hop] [pytest -k expression]
set -uo pipefail
cd "$(dirname "$0")/.."
tool=${1:-memcheck}
expr=${2:-"linear or attn or embed or sample or gemm"}
out=${SANITIZE_OUT:-gpurun_out}
mkdir -p "$out"
tools=("$tool")
[ "$tool" = all ] && tools=(memcheck racecheck synccheck)
if [ "$tool" = hop ]; then
  tools=(racecheck synccheck)
  expr=${2:-"hop or poison or third_units or half_blocks"}
  suffix=_hop
  files="tests/test_ops_gpu.py tests/test_engine_gpu.py"
fi
files=${files:-"tests/test_ops_gpu.py tests/test_gemm_gpu.py"}
suffix=${suffix:-}
rc=0
for t in "${tools[@]}"; do
  log="$out/sanitize_$t$suffix.log"
  echo "== compute-sanitizer --tool $t  (-k \"$expr\") -> $log"
  timeout 1500 compute-sanitizer --tool "$t" --error-exitcode 9 --launch-timeout 120 \
    python -m pytest $files -m gpu -x -q -k "$expr" -p no:cacheprovider > "$log" 2>&1
  r=$?
  tail -4 "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -2
  [ $r -ne 0 ] && rc=$r
done
exit $rc
