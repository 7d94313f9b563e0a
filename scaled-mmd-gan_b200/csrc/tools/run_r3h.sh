#!/bin/bash
# round 2, session 3, call H (1 GPU): evidence at HEAD -- GPU tests, smoke, bench line, C4 sweep, launch list of bench.py,
# ncu --set full of the bench-size kernels
mkdir -p gpurun_out
python -X faulthandler -m pytest tests -m gpu -q 2>&1 | grep -v "^  File \"/opt" | tail -6 | tee gpurun_out/r3h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r3h_smoke.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/r3h_bench.err | grep "^{" > gpurun_out/r3h_bench.json; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3h_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_sustained_peak'), d['e2e']['ms_per_step'], d['clocks'], d['kid']['ms_per_call'], d['kid']['e2e']['ms_per_call'])
print({k:(round(v['us_per_loss_stream'],1), round(v['us_per_loss_cuda_graph'],1)) for k,v in d['small_batch_latency'].items()})"
python bench.py --sweep --steps 10 2> gpurun_out/r3h_sweep.err | grep "^{" > gpurun_out/r3h_sweep.jsonl; echo "sweep rc=$?"
python -c "
import json
for l in open('gpurun_out/r3h_sweep.jsonl'):
    d=json.loads(l); print(d['n'],d['d'],'%.3f ms'%d['ms'],'%.0f TF'%d['tflops_algorithmic'],'%.3f'%d['frac_of_peak'],d['path'])"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3h_bench_plain.json 2> gpurun_out/r3h_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3h_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3h_ncu_bench.log 2>&1
echo "launch list exit=$?"
B=scaled-mmd-gan_b200/build/tc_check
export LD_LIBRARY_PATH=$PWD/scaled-mmd-gan_b200/lib
$B mmd mix_rq 65536 65536 256 2 0 > gpurun_out/r3h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_symf|tc_sym_wz" -s 2 -c 2 -o gpurun_out/r3h_symf_n65536 -f $B mmd mix_rq 65536 65536 256 2 0 > gpurun_out/r3h_ncu.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/r3h_ncu.log; cat gpurun_out/r3h_plain.log | grep -v "^   sum"
