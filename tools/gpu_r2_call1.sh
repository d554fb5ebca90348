#!/bin/bash
# round 2, GPU call 1: state of the suite with the new BASELINE-size parity tests, memcheck of the PCG
# tests, the slice-blocked SpMV micro-benchmark, Portfolio side by side with the reference CUDA build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c1_pytest.log 2>&1
( time timeout 300 ./tools/micro/spmv_sb ) > gpurun_out/r2c1_spmv_sb.log 2>&1
( time timeout 240 python tools/portfolio_study.py 100000 200 0.05 4000 ) > gpurun_out/r2c1_portfolio_mid.log 2>&1
( time timeout 420 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_kernels.py -q -k "pcg" --timeout 400 ) > gpurun_out/r2c1_memcheck_pcg.log 2>&1
tail -5 gpurun_out/r2c1_pytest.log
tail -30 gpurun_out/r2c1_spmv_sb.log
grep -c "ERROR SUMMARY" gpurun_out/r2c1_memcheck_pcg.log; grep "ERROR SUMMARY" gpurun_out/r2c1_memcheck_pcg.log | tail -3
