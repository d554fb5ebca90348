#!/bin/bash
# round 2, GPU call 5 (2 GPUs): the peer-memory CG loop for real (call 4 ran the NCCL path: the P2P branch of
# the solve was missing), lean passes with long-row chunks, sharded bench at scale 0.25 with / without P2P
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 -k "p2p or block" ) > gpurun_out/r2c5_pytest_sharded.log 2>&1
tail -12 gpurun_out/r2c5_pytest_sharded.log
( time timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_parity.py -q --timeout 600 -k "pcg or huber or svm or lasso_mid" ) > gpurun_out/r2c5_pytest_lean_long.log 2>&1
tail -5 gpurun_out/r2c5_pytest_lean_long.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 600 $TR bench.py --gpus 2 --scale 0.25 --steps 3 --warmup 2 ) > gpurun_out/r2c5_bench_p2p_s025.json 2> gpurun_out/r2c5_bench_p2p_s025_err.log
( time B200_DIST_NO_P2P=1 timeout 600 $TR bench.py --gpus 2 --scale 0.25 --steps 3 --warmup 2 --no-strong-baseline ) > gpurun_out/r2c5_bench_nccl_s025.json 2> gpurun_out/r2c5_bench_nccl_s025_err.log
python - <<'PY'
import json
for f in ("r2c5_bench_p2p_s025.json", "r2c5_bench_nccl_s025.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["gpu_launches"], d["status"], d["obj_val"], d["exchange"], d.get("strong_scaling"), d.get("roofline", {}).get("phases_us"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/r2c5_bench_p2p_s025_err.log
