#!/bin/bash
# round 2, GPU call 3 (1 GPU): batched small-QP kernel (tests + 4096 MPC QPs), SB v2 micro-benchmark with
# deeper pipelines, one short run of the new single-GPU bench line
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_batch.py -q --timeout 500 ) > gpurun_out/r2c3_pytest_batch.log 2>&1
( time timeout 300 python tools/batch_mpc.py 4096 ) > gpurun_out/r2c3_batch_mpc.log 2>&1
( time timeout 300 ./tools/micro/spmv_sb2 ) > gpurun_out/r2c3_spmv_sb2.log 2>&1
( time timeout 600 python bench.py --steps 3 --warmup 2 --same-config-budget 150 ) > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench_err.log
tail -15 gpurun_out/r2c3_pytest_batch.log
cat gpurun_out/r2c3_batch_mpc.log
cat gpurun_out/r2c3_spmv_sb2.log | cut -c1-300
tail -c 3000 gpurun_out/r2c3_bench.json; tail -5 gpurun_out/r2c3_bench_err.log
