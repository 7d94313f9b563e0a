#!/bin/bash
# round 2, call P: full GPU test suite + bench.py (mmd) + C4 sweep with the fused symmetric path as default
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2p_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2p_bench.err
python bench.py --sweep --steps 10 > gpurun_out/r2p_sweep.jsonl 2> gpurun_out/r2p_sweep.err; echo "sweep rc=$?"; tail -c 600 gpurun_out/r2p_sweep.err
cat gpurun_out/r2p_sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n'],d['d'],'%.3f ms'%d['ms'],'%.0f TF'%d['tflops_algorithmic'],'%.3f'%d['frac_of_peak'],d['path'])"
