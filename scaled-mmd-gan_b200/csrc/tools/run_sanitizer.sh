#!/bin/bash
# usage: run_sanitizer.sh <racecheck|synccheck|memcheck>   (one tool per gpurun call, B200_PROFILING.md)
# The suite runs every tensor-core gradient path once on ~1.5K-row problems inside ONE process.
mkdir -p gpurun_out
T=$1
B=scaled-mmd-gan_b200/build/tc_check
$B suite > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
cat gpurun_out/sanitizer_plain.log
timeout 900 compute-sanitizer --tool $T --print-limit 20 $B suite > gpurun_out/sanitizer_$T.log 2>&1
echo "compute-sanitizer --tool $T exit=$?" | tee -a gpurun_out/sanitizer_$T.log
tail -15 gpurun_out/sanitizer_$T.log
