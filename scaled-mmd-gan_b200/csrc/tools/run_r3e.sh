#!/bin/bash
# round 2, session 3, call E (N GPUs): peer-memory exchange vs NCCL on the bench workload, C5 latency, multi-GPU parity check
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --exchange peer 2> gpurun_out/r3e_bench_peer_n$N.err | grep "^{" > gpurun_out/r3e_bench_peer_n$N.json
timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl --mmd-only 2> gpurun_out/r3e_bench_nccl_n$N.err | grep "^{" > gpurun_out/r3e_bench_nccl_n$N.json
python - <<PY
import json
for k in ("peer", "nccl"):
    try:
        d = json.loads(open("gpurun_out/r3e_bench_%s_n$N.json" % k).read())
        print(k, "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
              "parity", (d.get("parity") or {}).get("grad_rel_to_max"), (d.get("parity") or {}).get("ok"), "kid", (d.get("kid") or {}).get("ms_per_call"))
    except Exception as e:
        print(k, "failed:", e)
PY
tail -3 gpurun_out/r3e_bench_peer_n$N.err
timeout 200 $TR --master-port 29523 bench_step.py --config c5 2> gpurun_out/r3e_c5_n$N.err | grep "^{" | tee gpurun_out/r3e_c5_n$N.json
tail -2 gpurun_out/r3e_c5_n$N.err
timeout 300 $TR --master-port 29524 tests/multi_gpu_check.py 2>&1 | grep "peers\|C5-size\|MULTI_GPU\|MISMATCH\|Error" | tee gpurun_out/r3e_multi_gpu_check_n$N.log
