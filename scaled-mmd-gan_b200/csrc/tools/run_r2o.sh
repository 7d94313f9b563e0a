#!/bin/bash
# round 2, call O: 112 registers per thread in the 576-thread kernels (no spills): fused pair, symmetric, fused symmetric
mkdir -p gpurun_out
L=gpurun_out/r2o.log
: > $L
B=scaled-mmd-gan_b200/build/tc_check
run() { echo "\$ $*  [SYM=$SMMD_SYM SYMF=$SMMD_SYMF MIN=$SMMD_SYM_MIN_ROWS]" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 1000 1100 256 2
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 1024 3 0
export SMMD_SYM=0
run $B mmd mix_rq 1000 1100 256 2
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
grep -vE "^   sum\[|^\[clock|value-only" $L
