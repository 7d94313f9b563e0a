#!/bin/bash
# Sweep the UMMA probe over candidate MN-major (LBO,SBO) encodings; each run in its own process + timeout.
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/umma_probe
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.log 2>&1
for mode in ts ss; do
  for cfg in "16384 1024" "1024 16384" "16384 2048" "2048 16384"; do
    timeout 30 $B $mode $cfg >> gpurun_out/probe.log 2>&1
    echo "exit=$?" >> gpurun_out/probe.log
  done
done
cat gpurun_out/probe.log
