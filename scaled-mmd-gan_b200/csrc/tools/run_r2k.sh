#!/bin/bash
# round 2, call K: fused symmetric variant (tc_symf_kernel + mirrored-only pass 2): parity on small forced shapes, timings
mkdir -p gpurun_out
L=gpurun_out/r2k.log
: > $L
B=scaled-mmd-gan_b200/build_dev/tc_check
run() { echo "\$ $*  [SYMF=$SMMD_SYMF MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rq 1000 1100 256 2
run $B mmd mix_rq_dot 900 1000 192 2
run $B mmd rbf 1024 1024 64 2
run $B mmd mix_rbf 2000 1500 64 2
run $B mmd distance 1500 1500 192 2
run $B mmd mix_rq 4096 4096 256 10
run $B mmd mix_rq 5000 3000 128 5
unset SMMD_SYM_MIN_ROWS
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd rbf 32768 32768 256 5 0
export SMMD_SYM_ONLY=1
run $B mmd mix_rq 32768 32768 256 5 0
export SMMD_SYM_ONLY=2
run $B mmd mix_rq 32768 32768 256 5 0
unset SMMD_SYM_ONLY
export SMMD_SYMF=0
run $B mmd mix_rq 32768 32768 256 5 0
grep -vE "^   sum\[|^\[clock|value-only" $L
