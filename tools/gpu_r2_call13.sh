#!/bin/bash
# round 2, GPU call 13 (1 GPU): GPU suite after the band fixes; sanitizers on the PCG tests with the loop body as
# plain launches (the sanitizer cannot follow kernel nodes of a conditional graph); initcheck
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c13_pytest.log 2>&1
tail -4 gpurun_out/r2c13_pytest.log
T="python -m pytest tests/test_gpu_kernels.py -q --timeout 500 -k pcg_against_direct_solve"
( B200_PCG_HOSTLOOP=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c13_memcheck_pcg_hostloop.log 2>&1
( B200_PCG_HOSTLOOP=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c13_racecheck_pcg_hostloop.log 2>&1
( B200_PCG_HOSTLOOP=1 timeout 400 compute-sanitizer --tool initcheck --print-limit 6 --error-exitcode 0 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve_parity.py -q --timeout 380 -k "pcg_against or large_qp or fused or transpose" ) > gpurun_out/r2c13_initcheck.log 2>&1
( timeout 500 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 python -m pytest tests/test_gpu_batch.py tests/test_gpu_kernels.py -q --timeout 480 -k "determinism or fused or residual" ) > gpurun_out/r2c13_racecheck_batch_fused.log 2>&1
for f in memcheck_pcg_hostloop racecheck_pcg_hostloop initcheck racecheck_batch_fused; do echo "== $f"; grep -v "Host Frame" gpurun_out/r2c13_$f.log | head -24 | cut -c1-240; grep -E "SUMMARY|passed|failed" gpurun_out/r2c13_$f.log | tail -3; done
