#!/bin/bash
# round 2, GPU call 24 (2 GPUs): the driver's launch line at N = 2 after the store_solution change
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562"
( time timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2c24_bench_svm_2gpu.json 2> gpurun_out/r2c24_bench_svm_2gpu_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c24_bench_svm_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["e2e"], d["exchange"], d.get("strong_scaling"))
PY
tail -n 4 gpurun_out/r2c24_bench_svm_2gpu_err.log
