#!/usr/bin/env bash
# 8-GPU measurement set (run under `gpurun --gpus 8`): results in gpurun_out/
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
B="bench.py --gpus 8 --steps 64 --warmup 8"
show() { python -c "import sys,json; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('config', {}).get('parallelism'), 'wait', d.get('stage_wait_us_per_step'), 'busy', d.get('stage_busy_us_per_step'))" $1; }
python -m pytest tests/test_multiprocess_gpu.py -m gpu -x -q -k "device_mode_matches_single_gpu and 8" 2>&1 | tail -3
$TR --master-port 29521 $B --partition half > gpurun_out/bench_n8_half.json 2> gpurun_out/bench_n8.err; show gpurun_out/bench_n8_half.json
$TR --master-port 29522 $B --partition balanced > gpurun_out/bench_n8.json 2>> gpurun_out/bench_n8.err; show gpurun_out/bench_n8.json
$TR --master-port 29523 $B --partition half --weights fp8 --prompt-len 1024 --seq-len 2048 > gpurun_out/bench_n8_fp8_2048ctx.json 2>> gpurun_out/bench_n8.err; show gpurun_out/bench_n8_fp8_2048ctx.json
$TR --master-port 29524 $B --partition half --hop nccl > gpurun_out/bench_n8_nccl.json 2>> gpurun_out/bench_n8.err; show gpurun_out/bench_n8_nccl.json
tail -5 gpurun_out/bench_n8.err
