#!/bin/bash
# round 2, call T: full GPU test suite, smoke(), bench.py with the C++ wrapper (small-batch latency), bench_step C2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/r2t_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
SMMD_NO_EXT=1 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2t_bench_noext.json 2> gpurun_out/r2t_bench_noext.err; echo "bench(no ext) rc=$?"
python bench_step.py --steps 30 --warmup 5 > gpurun_out/r2t_c2.json 2> gpurun_out/r2t_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2t_bench.json","gpurun_out/r2t_bench_noext.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "ms/step %.2f frac %.3f e2e %.2f  each %s" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("ms_each_step")))
    print("   small:", {k:(round(v["us_per_loss_stream"],1), round(v["us_per_loss_cuda_graph"],1)) for k,v in d["small_batch_latency"].items()})
d=json.loads([l for l in open("gpurun_out/r2t_c2.json") if l.startswith("{")][-1])
for k in ("yml_rbf_dof1","mix_rbf_dof16"):
    print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in d[k].items()})
PY
