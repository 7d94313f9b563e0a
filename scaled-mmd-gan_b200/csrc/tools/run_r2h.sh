#!/bin/bash
# round 2, call H: sleeping waits in the fused / two-pass / Gram kernels (regression + gain), small-N symmetric cells,
# GPU test suite, ncu --set full of the symmetric pass-1 kernel
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
B=scaled-mmd-gan_b200/build/tc_check
run() { echo "\$ $*  [SYM=$SMMD_SYM MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM=0
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 4096 4096 512 20 0
run $B mmd mix_rq 4096 4096 1024 20 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B kid 50000 2048 100 1000 3
unset SMMD_SYM
export SMMD_SYM_MIN_ROWS=1
run $B mmd mix_rq 4096 4096 512 20 0
run $B mmd mix_rq 4096 4096 1024 20 0
run $B mmd mix_rq 4096 4096 256 20 0
unset SMMD_SYM_MIN_ROWS
grep -vE "^   sum\[|^\[clock|value-only|dX:|dY:" $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2h_pytest.log
$B mmd mix_rq 16384 16384 512 2 0 > gpurun_out/r2h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_sym_wgen -s 1 -c 1 -o gpurun_out/r2h_sym -f $B mmd mix_rq 16384 16384 512 2 0 > gpurun_out/r2h_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2h_ncu.log
