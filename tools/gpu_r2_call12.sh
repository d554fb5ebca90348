#!/bin/bash
# diagnostic: memcheck / racecheck attribution for test_large_qp[graph]
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_solve_parity.py -q -x --timeout 300 -k large_qp"
( timeout 300 compute-sanitizer --tool memcheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c12_memcheck_graph.log 2>&1
( B200_PCG_HOSTLOOP=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c12_memcheck_hostloop.log 2>&1
( B200_PCG_HOSTLOOP=1 timeout 300 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c12_racecheck_hostloop.log 2>&1
( B200_PCG_HOSTLOOP=1 timeout 300 compute-sanitizer --tool initcheck --print-limit 6 --error-exitcode 0 $T ) > gpurun_out/r2c12_initcheck_hostloop.log 2>&1
for f in memcheck_graph memcheck_hostloop racecheck_hostloop initcheck_hostloop; do echo "== $f"; grep -v "Host Frame" gpurun_out/r2c12_$f.log | head -40 | cut -c1-260; done
