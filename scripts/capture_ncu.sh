#!/usr/bin/env bash
# ncu --set full captures of the three hot kernels (one GPU; never a bench value).  Reports land in
# gpurun_out/prof_<what>.ncu-rep; summarise here with scripts/summarise_ncu.sh into profiles/.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for what in ${@:-decode gemm attn}; do
  case $what in
    decode) k="regex:stream_bulk_kernel" ;;
    gemm)   k="regex:gemm_bf16_tcgen05_pair_kernel" ;;
    attn)   k="regex:attn_decode_kernel" ;;
    attnp)  k="regex:attn_prefill_tcgen05_kernel" ;;
    gemm_fp8) k="regex:gemm_fp8_blockscaled_kernel" ;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "$k" --launch-skip 2 --launch-count 1 -f \
    -o gpurun_out/prof_$what python scripts/profile_targets.py $what > gpurun_out/ncu_$what.log 2>&1
  tail -2 gpurun_out/ncu_$what.log
done
