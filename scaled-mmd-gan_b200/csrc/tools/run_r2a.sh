#!/bin/bash
# round 2, call A: A-operand MN-major probe + baseline timings of the round-1 kernels
mkdir -p gpurun_out
L=gpurun_out/r2a.log
: > $L
P=scaled-mmd-gan_b200/build/umma3_probe
for cfg in "8192 1024 2048" "1024 8192 2048" "8192 2048 2048" "16 1024 2048" "8192 1024 32"; do
  timeout 30 $P $cfg >> $L 2>&1; echo "exit=$?" >> $L
done
B=scaled-mmd-gan_b200/build/tc_check
run() { echo "\$ $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $B mmd mix_rq 4096 4096 256 20 0
run $B mmd mix_rq 8192 8192 256 20 0
run $B mmd mix_rq 16384 16384 256 10 0
run $B mmd mix_rq 32768 32768 256 5 0
run $B mmd mix_rq 8192 8192 512 10 0
run $B mmd mix_rq 32768 32768 512 3 0
run $B mmd mix_rq 8192 8192 1024 10 0
run $B mmd mix_rq 32768 32768 1024 3 0
run $B mmd rbf 32768 32768 256 5 0
grep -vE "^   sum\[" $L
