#!/bin/bash
# round 2, call F: ncu --set full of the symmetric pass-1 kernel v3 at N = 16384+16384, d = 256
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
$B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_sym_wgen -s 1 -c 1 -o gpurun_out/r2f_sym -f $B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2f_ncu.log 2>&1
echo "ncu exit=$?" >> gpurun_out/r2f_ncu.log
tail -3 gpurun_out/r2f_ncu.log
