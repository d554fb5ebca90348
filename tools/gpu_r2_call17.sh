#!/bin/bash
# round 2, GPU call 17 (4 GPUs): the driver's own launch line at N = 4 (full-size configs[3], --steps 20 --warmup 5)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 800 $TR bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r2c17_bench_svm_4gpu.json 2> gpurun_out/r2c17_bench_svm_4gpu_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c17_bench_svm_4gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["e2e"], d["exchange"], d.get("strong_scaling"), d.get("roofline", {}).get("phases_us"))
PY
tail -n 4 gpurun_out/r2c17_bench_svm_4gpu_err.log
