#!/bin/bash
# round 2, call C: symmetric path with barrier-free per-warp epilogue: parity + timing + per-pass breakdown
mkdir -p gpurun_out
L=gpurun_out/r2c.log
: > $L
B=scaled-mmd-gan_b200/build/tc_check
run() { echo "\$ $*  [MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rq 1000 1100 256 2
run $B mmd rbf 1024 1024 512 2
run $B mmd mix_rq 4096 4096 256 20
unset SMMD_SYM_MIN_ROWS
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 1024 3 0
for o in 1 2; do
export SMMD_SYM_ONLY=$o
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B mmd rbf 32768 32768 256 5 0
done
grep -vE "^   sum\[|^\[clock|value-only" $L
