#!/bin/bash
# round 2, call D: ncu --set full of the symmetric pass-1 kernel (and pass 2) at N = 8192+8192, d = 256
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
$B mmd mix_rq 8192 8192 256 3 0 > gpurun_out/r2d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_sym -s 2 -c 2 -o gpurun_out/r2d_sym -f $B mmd mix_rq 8192 8192 256 3 0 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu exit=$?" >> gpurun_out/r2d_ncu.log
tail -5 gpurun_out/r2d_ncu.log
