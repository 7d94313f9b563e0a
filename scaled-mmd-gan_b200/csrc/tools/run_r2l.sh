#!/bin/bash
# round 2, call L: per-pass times of the fused symmetric variant; ncu --set full of tc_symf_kernel
mkdir -p gpurun_out
L=gpurun_out/r2l.log
: > $L
B=scaled-mmd-gan_b200/build_dev/tc_check
run() { echo "\$ $*  [SYMF=$SMMD_SYMF MIN=$SMMD_SYM_MIN_ROWS ONLY=$SMMD_SYM_ONLY]" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export SMMD_SYM_MIN_ROWS=1
for n in 8192 32768; do
export SMMD_SYM_ONLY=1
run $B mmd mix_rq $n $n 256 5 0
export SMMD_SYM_ONLY=2
run $B mmd mix_rq $n $n 256 5 0
unset SMMD_SYM_ONLY
done
grep -vE "^   sum\[|^\[clock|value-only" $L
$B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2l_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_symf|tc_sym_wz" -s 2 -c 2 -o gpurun_out/r2l_symf -f $B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2l_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2l_ncu.log
