#!/bin/bash
# round 2, session 3, call I: device-counted steps (1 GPU tests of the peers entry; N GPUs: parity check + C5 incl. graph replay)
mkdir -p gpurun_out
N=${1:-2}
SKIP1=${2:-}
[ -z "$SKIP1" ] && python -X faulthandler -m pytest tests/test_gpu_peers.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r3i_pytest_peers.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29554 tests/multi_gpu_check.py 2>&1 | grep "C5-size\|MULTI_GPU\|MISMATCH\|Error" | tee gpurun_out/r3i_multi_gpu_check_n$N.log
timeout 200 $TR --master-port 29553 bench_step.py --config c5 2> gpurun_out/r3i_c5_n$N.err | grep "^{" | tee gpurun_out/r3i_c5_n$N.json
tail -3 gpurun_out/r3i_c5_n$N.err
