#!/bin/bash
# round 2, call J: GPU tests (full), bench.py (mmd + kid workloads), bench_step C2, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2j_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2j_bench.err
python bench.py --workload kid --steps 8 > gpurun_out/r2j_kid.json 2> gpurun_out/r2j_kid.err; echo "kid rc=$?"; tail -c 1500 gpurun_out/r2j_kid.err
python bench_step.py --steps 30 --warmup 5 > gpurun_out/r2j_c2.json 2> gpurun_out/r2j_c2.err; echo "c2 rc=$?"; tail -c 1500 gpurun_out/r2j_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; echo "ref rc=$?"; tail -c 1000 gpurun_out/r2j_ref.err
nproc; free -g | head -2
