#!/bin/bash
mkdir -p gpurun_out
D=scaled-mmd-gan_b200/build/tc_check_dbg
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/abl.log
: > $L
run() { echo "\$ $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for K in 1 2; do
export SMMD_FUSED_KSPLIT=$K
echo "=== KSPLIT=$K correctness" >> $L
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rbf 1000 1000 128 3
run $B mmd distance 512 512 192 3
run $B mmd mix_rq 4096 4096 256 20
echo "=== KSPLIT=$K timing (instrumented)" >> $L
SMMD_DEBUG_NULLMATH=1 run $D mmd mix_rq 8192 8192 256 10 0
run $D mmd mix_rq 8192 8192 256 10 0
run $D mmd rbf 8192 8192 256 10 0
echo "=== KSPLIT=$K perf" >> $L
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd rbf 8192 8192 256 20 0
run $B mmd mix_rbf 8192 8192 256 20 0
run $B mmd mix_rq 32768 32768 256 5 0
done
grep -E "===|TC  path|pipe timing|rel diff|dX|error|exit=[1-9]" $L
