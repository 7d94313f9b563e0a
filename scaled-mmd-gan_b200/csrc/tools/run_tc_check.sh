#!/bin/bash
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/tc_check.log
: > $L
run() { echo "\$ $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $B mmd mix_rq 300 200 100 1
run $B mmd mix_rbf 1000 1000 128 3
run $B mmd distance 512 512 192 3
run $B mmd rbf 700 900 256 3
run $B mmd mix_rq 4096 4096 256 20
grep -vE "^   sum\[" $L
