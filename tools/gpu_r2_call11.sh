#!/bin/bash
# round 2, GPU call 11 (8 GPUs): the north-star configuration -- ONE configs[3] QP row-sharded over 8 B200
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 300 $TR tools/sharded_worker.py --family svm --scale 0.003 --check --blocks ) > gpurun_out/r2c11_worker_svm8.log 2>&1
grep -E "SHARDED|REPLICATED" gpurun_out/r2c11_worker_svm8.log | cut -c1-700
( time timeout 800 $TR bench.py --gpus 8 --steps 5 --warmup 2 ) > gpurun_out/r2c11_bench_svm_8gpu.json 2> gpurun_out/r2c11_bench_svm_8gpu_err.log
( time timeout 200 $TR bench.py --gpus 8 --steps 5 --warmup 2 --mode batch --workload mpc ) > gpurun_out/r2c11_bench_mpc_batch_8gpu.json 2> gpurun_out/r2c11_bench_mpc_batch_8gpu_err.log
cat gpurun_out/r2c11_bench_mpc_batch_8gpu.json | cut -c1-1500
( time timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 2 --workload huber --no-strong-baseline ) > gpurun_out/r2c11_bench_huber_8gpu.json 2> gpurun_out/r2c11_bench_huber_8gpu_err.log
python - <<'PY'
import json
for f in ("r2c11_bench_svm_8gpu.json", "r2c11_bench_huber_8gpu.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["e2e"], d["exchange"], d.get("strong_scaling"), d.get("roofline", {}).get("phases_us"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/r2c11_bench_svm_8gpu_err.log
