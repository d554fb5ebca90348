#!/bin/bash
# round 2, GPU call 10 (2 GPUs): sharded parity after the termination-check reductions and vector heads moved to
# the peer-memory path as well
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 ) > gpurun_out/r2c10_pytest_sharded.log 2>&1
tail -8 gpurun_out/r2c10_pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 600 $TR bench.py --gpus 2 --scale 0.25 --steps 3 --warmup 2 --no-strong-baseline ) > gpurun_out/r2c10_bench_p2p_s025.json 2> gpurun_out/r2c10_bench_p2p_s025_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c10_bench_p2p_s025.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["exchange"])
PY
tail -3 gpurun_out/r2c10_bench_p2p_s025_err.log
