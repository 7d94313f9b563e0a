#!/bin/bash
# hardware check of the cluster (d > 256) fused kernel: parity vs the exact path at small sizes, then timing
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
L=gpurun_out/cluster_check.log
: > $L
run() { echo "\$ $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $B mmd mix_rq 300 200 512 1
run $B mmd mix_rq 1024 1024 512 2
run $B mmd mix_rq 1000 1100 384 2
run $B mmd rbf 1024 1024 512 2
run $B mmd mix_rq 1024 1024 1024 2
run $B mmd mix_rq 900 1000 768 2
run $B mmd mix_rq 4096 4096 512 10
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 8192 8192 1024 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B mmd rbf 32768 32768 512 3 0
grep -vE "^   sum\[" $L
