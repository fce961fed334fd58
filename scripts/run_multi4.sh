TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_multiprocess_gpu.py -m gpu -x -q 2>&1 | tail -3
$TR --nproc-per-node 4 --master-port 29511 bench.py --gpus 4 --steps 64 --warmup 8 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_n4.json
$TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --steps 64 --warmup 8 --hop nccl > gpurun_out/bench_n4_nccl.json 2>> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_n4_nccl.json
$TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench_n2.json 2>> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_n2.json
$TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 64 --warmup 8 --hop nccl > gpurun_out/bench_n2_nccl.json 2>> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_n2_nccl.json
$TR --nproc-per-node 4 --master-port 29515 bench.py --gpus 4 --steps 64 --warmup 8 --model tiny-llama-1.1b > gpurun_out/bench_tinyllama_n4.json 2>> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_tinyllama_n4.json
$TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 --steps 64 --warmup 8 --weights fp8 > gpurun_out/bench_n4_fp8.json 2>> gpurun_out/bench_n4.err; cut -c1-160 gpurun_out/bench_n4_fp8.json
timeout 300 $TR --nproc-per-node 4 --master-port 29517 bench.py --impl reference --gpus 4 --steps 32 --warmup 4 --model tiny-llama-1.1b > gpurun_out/bench_ref_tiny_n4.json 2> gpurun_out/bench_ref_n4.err; cut -c1-200 gpurun_out/bench_ref_tiny_n4.json
tail -5 gpurun_out/bench_n4.err
