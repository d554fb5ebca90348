#!/bin/bash
# round 2, GPU call 6 (2 GPUs): full-size configs[3] SVM (1.2e8 nnz) row-sharded over 2 ranks, after the
# per-vector shard flags went in; quick parity check of the sharded path first
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 300 $TR tools/sharded_worker.py --family svm --scale 0.003 --check --blocks ) > gpurun_out/r2c6_worker_svm.log 2>&1
grep -E "SHARDED|REPLICATED" gpurun_out/r2c6_worker_svm.log | cut -c1-700
( time timeout 300 $TR tools/sharded_worker.py --family lasso --scale 0.003 --check ) > gpurun_out/r2c6_worker_lasso.log 2>&1
grep -E "SHARDED|REPLICATED" gpurun_out/r2c6_worker_lasso.log | cut -c1-700
( time timeout 800 $TR bench.py --gpus 2 --steps 5 --warmup 2 ) > gpurun_out/r2c6_bench_svm_2gpu.json 2> gpurun_out/r2c6_bench_svm_2gpu_err.log
python - <<'PY'
import json
for f in ("r2c6_bench_svm_2gpu.json",):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["e2e"], d["exchange"], d.get("strong_scaling"), d.get("roofline", {}).get("phases_us"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/r2c6_bench_svm_2gpu_err.log
