#!/bin/bash
# hardware check of the two-pass (W panel + GEMM) path: parity vs the exact path, multi-panel, timing
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/wz_check.log
: > $L
run() { echo "\$ $*  [MIN_D=$SMMD_WZ_MIN_D PANEL_MB=$SMMD_WZ_PANEL_MB]" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $B mmd mix_rq 300 200 1024 1
run $B mmd mix_rq 1000 1100 600 2
run $B mmd mix_rq_dot 900 1000 768 2
run $B mmd rbf 1024 1024 2048 2
export SMMD_WZ_PANEL_MB=1
run $B mmd mix_rq 1000 1100 600 2
unset SMMD_WZ_PANEL_MB
export SMMD_WZ_MIN_D=64
run $B mmd mix_rq 1000 1100 200 2
run $B mmd mix_rq 4096 4096 512 10
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 256 3 0
unset SMMD_WZ_MIN_D
run $B mmd mix_rq 4096 4096 1024 10 0
run $B mmd mix_rq 8192 8192 1024 10 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B mmd rbf 32768 32768 1024 3 0
run $B mmd mix_rq 65536 65536 1024 2 0
grep -vE "^   sum\[" $L
