#!/bin/bash
# round 2, session 3, call F (N GPUs): pipelined e2e with copies ordered behind the NVLink pull; parity check
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --exchange peer ${2:---mmd-only} 2> gpurun_out/r3f_bench_peer_n$N.err | grep "^{" > gpurun_out/r3f_bench_peer_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r3f_bench_peer_n$N.json").read())
print("peer ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
      "parity", (d.get("parity") or {}).get("grad_rel_to_max"), (d.get("parity") or {}).get("ok"), "kid", (d.get("kid") or {}).get("ms_per_call"))
PY
tail -3 gpurun_out/r3f_bench_peer_n$N.err
timeout 300 $TR --master-port 29534 tests/multi_gpu_check.py 2>&1 | grep "C5-size\|MULTI_GPU\|MISMATCH\|Error" | tee gpurun_out/r3f_multi_gpu_check_n$N.log
