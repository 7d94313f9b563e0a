#!/bin/bash
# round 2, session 3, call B: launch lists (ncu, cold-cache per-launch times) of the mid-size problems
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/scaled-mmd-gan_b200/lib
B=scaled-mmd-gan_b200/build/tc_check
for cfg in "8192 256" "16384 256" "8192 512"; do
set -- $cfg
timeout 300 $B mmd mix_rq $1 $1 $2 20 0 2>&1 | grep -v "sum\[" | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3b_launches_$1x$2.csv $B mmd mix_rq $1 $1 $2 4 0 > /dev/null 2>&1
done
