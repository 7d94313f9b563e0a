#!/bin/bash
# round 2, session 3, call D (8 GPUs): e2e breakdown of the host-buffer loop, with and without NUMA binding
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r3d_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r3d_topo.txt
N=${1:-8}
IFS=","; for flag in ${FLAGS:-,--bind}; do IFS=" "
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scaled-mmd-gan_b200/csrc/tools/e2e_probe.py $flag 2>&1 | grep "E2E_PROBE\|TL \|rank .* gpu\|Error\|error" | tee -a gpurun_out/r3d_probe.log
done
