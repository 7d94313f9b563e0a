#!/bin/bash
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/tc_check.log
: > $L
run() { echo "\$ $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for K in 1 2; do
export SMMD_FUSED_KSPLIT=$K
echo "=== KSPLIT=$K" >> $L
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rq 4096 4096 256 20
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd rbf 8192 8192 256 20 0
run $B mmd mix_rbf 8192 8192 256 20 0
run $B mmd mix_rq 32768 32768 256 5 0
done
grep -vE "^   sum\[" $L
