#!/bin/bash
# round 2, call G: sleeping mbarrier waits; split (software-pipelined) vs plain pass-1 epilogue; regression check of the other paths
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
B=scaled-mmd-gan_b200/build/tc_check
BN=scaled-mmd-gan_b200/build/tc_check_nosplit
run() { echo "\$ $*  [SYM=$SMMD_SYM MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 1000 1100 256 2
run $BN mmd mix_rq 1000 1100 256 2
run $B mmd mix_rbf 2000 1500 64 2
unset SMMD_SYM_MIN_ROWS
for b in $B $BN; do
run $b mmd mix_rq 8192 8192 256 20 0
run $b mmd mix_rq 16384 16384 256 10 0
run $b mmd mix_rq 32768 32768 256 5 0
done
run $B mmd mix_rq 65536 65536 256 3 0
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 8192 8192 1024 10 0
run $B mmd mix_rq 32768 32768 1024 3 0
export SMMD_SYM_ONLY=1
run $B mmd mix_rq 32768 32768 256 5 0
run $BN mmd mix_rq 32768 32768 256 5 0
export SMMD_SYM_ONLY=2
run $B mmd mix_rq 32768 32768 256 5 0
unset SMMD_SYM_ONLY
export SMMD_SYM=0
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B kid 50000 2048 100 1000 3
grep -vE "^   sum\[|^\[clock|value-only" $L
