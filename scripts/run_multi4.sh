#!/usr/bin/env bash
# 4-GPU measurement set (run under `gpurun --gpus 4`): results in gpurun_out/
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 64 --warmup 8"
show() { python -c "import sys,json; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('config', {}).get('parallelism'), 'wait', d.get('stage_wait_us_per_step'), 'busy', d.get('stage_busy_us_per_step'))" $1; }
python -m pytest tests/test_multiprocess_gpu.py -m gpu -x -q -k "half or device_mode_matches_single_gpu and 4" 2>&1 | tail -3
$TR --nproc-per-node 4 --master-port 29511 $B --gpus 4 --partition balanced > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; show gpurun_out/bench_n4.json
$TR --nproc-per-node 4 --master-port 29512 $B --gpus 4 --partition half > gpurun_out/bench_n4_half.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_n4_half.json
$TR --nproc-per-node 4 --master-port 29513 $B --gpus 4 --partition balanced --hop nccl > gpurun_out/bench_n4_nccl.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_n4_nccl.json
$TR --nproc-per-node 4 --master-port 29514 $B --gpus 4 --partition half --weights fp8 > gpurun_out/bench_n4_fp8.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_n4_fp8.json
$TR --nproc-per-node 4 --master-port 29515 $B --gpus 4 --model tiny-llama-1.1b > gpurun_out/bench_tinyllama_n4.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_tinyllama_n4.json
$TR --nproc-per-node 2 --master-port 29516 $B --gpus 2 > gpurun_out/bench_n2.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_n2.json
$TR --nproc-per-node 2 --master-port 29517 $B --gpus 2 --hop nccl > gpurun_out/bench_n2_nccl.json 2>> gpurun_out/bench_n4.err; show gpurun_out/bench_n2_nccl.json
tail -5 gpurun_out/bench_n4.err
