#!/bin/bash
# round 2, call S (8 GPUs): sharded-vs-single parity, bench.py headline + KID workload, C5 loss latency, C4 cells at N = 65536
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29701 tests/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU|MISMATCH" | tee gpurun_out/r2s_check.log
$TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err; echo "bench rc=$?"
$TR --master-port 29703 bench.py --gpus 8 --workload kid --steps 10 > gpurun_out/r2s_kid_n8.json 2> gpurun_out/r2s_kid_n8.err; echo "kid rc=$?"
$TR --master-port 29704 bench_step.py --config c5 > gpurun_out/r2s_c5.json 2> gpurun_out/r2s_c5.err; echo "c5 rc=$?"; tail -c 600 gpurun_out/r2s_c5.err
$TR --master-port 29705 bench.py --gpus 8 --sweep --sweep-n 65536 --steps 10 > gpurun_out/r2s_sweep_n8.jsonl 2> gpurun_out/r2s_sweep_n8.err; echo "sweep rc=$?"
tail -c 400 gpurun_out/r2s_bench_n8.err
