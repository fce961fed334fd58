#!/usr/bin/env bash
# 2-GPU verification set (run under `gpurun --gpus 2`)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 python -m pytest tests/test_multiprocess_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "2 or half or multi_gpu" 2>&1 | tail -3
timeout 300 $TR --master-port 29531 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-160 gpurun_out/bench_n2.json
timeout 500 $TR --master-port 29532 bench.py --impl reference --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cut -c1-300 gpurun_out/bench_ref_n2.json; tail -2 gpurun_out/bench_ref_n2.err | cut -c1-200
