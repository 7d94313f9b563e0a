#!/bin/bash
# round 2, call M: fused symmetric variant v2 (norm ring in shared memory, longer units / pieces): parity + timings
mkdir -p gpurun_out
L=gpurun_out/r2m.log
: > $L
B=scaled-mmd-gan_b200/build_dev/tc_check
run() { echo "\$ $*  [SYMF=$SMMD_SYMF MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rq 1000 1100 256 2
run $B mmd mix_rbf 2000 1500 64 2
run $B mmd distance 1500 1500 192 2
run $B mmd mix_rq 5000 3000 128 5
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd rbf 32768 32768 256 5 0
for n in 16384 32768; do
export SMMD_SYM_ONLY=1
run $B mmd mix_rq $n $n 256 5 0
export SMMD_SYM_ONLY=2
run $B mmd mix_rq $n $n 256 5 0
unset SMMD_SYM_ONLY
done
grep -vE "^   sum\[|^\[clock|value-only" $L
