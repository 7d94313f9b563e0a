#!/bin/bash
# round 2, call N: ncu --set full of tc_symf_kernel v2 + mirrored pass at N = 16384 + 16384, d = 256
mkdir -p gpurun_out
B=scaled-mmd-gan_b200/build/tc_check
export SMMD_SYM_MIN_ROWS=1
$B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2n_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_symf|tc_sym_wz" -s 2 -c 2 -o gpurun_out/r2n_symf -f $B mmd mix_rq 16384 16384 256 2 0 > gpurun_out/r2n_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2n_ncu.log
