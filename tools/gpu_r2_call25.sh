#!/bin/bash
# round 2, GPU call 25 (4 GPUs): the driver's launch line at N = 4 after the store_solution change
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563"
( time timeout 200 $TR bench.py --gpus 4 --steps 10 --warmup 3 ) > gpurun_out/r2c25_bench_svm_4gpu.json 2> gpurun_out/r2c25_bench_svm_4gpu_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c25_bench_svm_4gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["e2e"]["time_to_solution_ms"], d["exchange"]["nccl_allreduce_calls_per_solve"], d.get("strong_scaling"))
PY
tail -n 4 gpurun_out/r2c25_bench_svm_4gpu_err.log
