#!/bin/bash
# round 2, call E: symmetric path v3 (compact units, smem norms, pair terms): parity + timing + per-pass breakdown
mkdir -p gpurun_out
L=gpurun_out/r2e.log
: > $L
B=scaled-mmd-gan_b200/build/tc_check
run() { echo "\$ $*  [MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rq 1000 1100 256 2
run $B mmd mix_rq_dot 900 1000 320 2 
run $B mmd rbf 1024 1024 512 2
run $B mmd mix_rbf 2000 1500 64 2
run $B mmd distance 1500 1500 192 2
run $B mmd mix_rq 4096 4096 256 20
run $B mmd mix_rq 5000 3000 1024 5
unset SMMD_SYM_MIN_ROWS
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 8192 8192 1024 10 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B mmd mix_rq 65536 65536 1024 2 0
export SMMD_SYM_ONLY=1
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd rbf 32768 32768 256 5 0
grep -vE "^   sum\[|^\[clock|value-only" $L
