#!/bin/bash
# round 2, GPU call 2 (1 GPU): TMA-fed slice-blocked SpMV micro-benchmark; GPU suite on the rebuilt libraries
mkdir -p gpurun_out
( time timeout 300 ./tools/micro/spmv_sb2 ) > gpurun_out/r2c2_spmv_sb2.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c2_pytest.log 2>&1
cat gpurun_out/r2c2_spmv_sb2.log
tail -5 gpurun_out/r2c2_pytest.log
