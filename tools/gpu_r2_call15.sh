#!/bin/bash
# round 2, GPU call 15 (1 GPU): batched kernel at 4 / 5 / 6 CTAs per SM, full GPU suite on the final libraries, setup trace
mkdir -p gpurun_out
for v in "" build/var5 build/var6; do
  echo "== B200_LIBDIR=$v"
  ( B200_LIBDIR=$v timeout 200 python tools/batch_mpc.py 4096 --cpu-sample 0 ) 2>&1 | tail -1 | cut -c1-420
  ( B200_LIBDIR=$v timeout 200 python tools/batch_mpc.py 16384 --cpu-sample 0 ) 2>&1 | tail -1 | cut -c1-420
done > gpurun_out/r2c15_batch_variants.log 2>&1
cat gpurun_out/r2c15_batch_variants.log
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c15_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2c15_pytest.log | tail -2; grep -E "^FAILED" gpurun_out/r2c15_pytest.log | head
( B200_TRACE_SETUP=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2c15_bench_trace.json 2> gpurun_out/r2c15_setup_trace.log
grep "b200 trace" gpurun_out/r2c15_setup_trace.log | tail -24
