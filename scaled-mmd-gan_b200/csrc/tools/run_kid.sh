#!/bin/bash
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/kid.log
: > $L
run() { echo "\$ $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $B kid 2000 256 4 250 3
run $B kid 3000 300 5 700 3
run $B kid 5000 2048 10 1000 10
run $B kid 20000 2048 100 1000 5 0
cat $L
