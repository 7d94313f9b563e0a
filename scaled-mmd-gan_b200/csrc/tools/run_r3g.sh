#!/bin/bash
# round 2, session 3, call G (8 GPUs): headline line with the peer exchange + pipelined e2e, C5 latency
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r3g_bench_n$N.err | grep "^{" > gpurun_out/r3g_bench_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r3g_bench_n$N.json").read())
print("peer ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
      "parity", (d.get("parity") or {}).get("grad_rel_to_max"), (d.get("parity") or {}).get("ok"), "kid", (d.get("kid") or {}).get("ms_per_call"))
PY
tail -2 gpurun_out/r3g_bench_n$N.err
timeout 200 $TR --master-port 29543 bench_step.py --config c5 2> gpurun_out/r3g_c5_n$N.err | grep "^{" | tee gpurun_out/r3g_c5_n$N.json
timeout 200 $TR --master-port 29544 bench.py --gpus $N --workload kid --steps 20 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r3g_bench_kid_n$N.json
