#!/bin/bash
# round 2, call Y: GPU tests, smoke, C4 sweep and bench with the final dispatch (sym from 4096 rows, symf from 40960 at d <= 256)
mkdir -p gpurun_out
python -X faulthandler -m pytest tests -m gpu -q 2>&1 | grep -v "^  File \"/opt" | tail -10 | tee gpurun_out/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --sweep --steps 10 2> gpurun_out/r2y_sweep.err | grep "^{" > gpurun_out/r2y_sweep.jsonl; echo "sweep rc=$?"
python -c "
import sys,json
for l in open('gpurun_out/r2y_sweep.jsonl'):
    d=json.loads(l); print(d['n'],d['d'],'%.3f ms'%d['ms'],'%.0f TF'%d['tflops_algorithmic'],'%.3f'%d['frac_of_peak'],d['path'])"
python bench.py --steps 20 --warmup 5 2> gpurun_out/r2y_bench.err | grep "^{" > gpurun_out/r2y_bench.json; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read())
print(d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_sustained_peak'), d['e2e']['ms_per_step'], d['clocks'], d['kid']['ms_per_call'], d['kid']['e2e']['ms_per_call'])
print({k:(round(v['us_per_loss_stream'],1), round(v['us_per_loss_cuda_graph'],1)) for k,v in d['small_batch_latency'].items()})"
