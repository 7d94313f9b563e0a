#!/bin/bash
# round 2, call R: evidence at HEAD -- launch list of the bench.py command, ncu --set full of the bench-size kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2r_bench_plain.json 2> gpurun_out/r2r_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2r_ncu_bench.log 2>&1
echo "launch list exit=$?"
B=scaled-mmd-gan_b200/build/tc_check
$B mmd mix_rq 65536 65536 256 2 0 > gpurun_out/r2r_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_symf|tc_sym_wz" -s 2 -c 2 -o gpurun_out/r2r_symf_n65536 -f $B mmd mix_rq 65536 65536 256 2 0 > gpurun_out/r2r_ncu.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/r2r_ncu.log; cat gpurun_out/r2r_plain.log | grep -v "^   sum"
