#!/bin/bash
# round 2, GPU call 14 (1 GPU): GPU suite after the schedule change (rows > half a tile -> CTA-wide chunk)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c14_pytest.log 2>&1
tail -4 gpurun_out/r2c14_pytest.log
grep -E "^FAILED|^E  " gpurun_out/r2c14_pytest.log | head -10 | cut -c1-300
( time B200_PCG_HOSTLOOP=1 timeout 240 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 python tools/racecheck_target.py ) > gpurun_out/r2c14_racecheck_target.log 2>&1
grep -v "Host Frame" gpurun_out/r2c14_racecheck_target.log | head -16 | cut -c1-240
( time B200_PCG_HOSTLOOP=1 timeout 120 compute-sanitizer --tool memcheck --print-limit 6 --error-exitcode 0 python tools/racecheck_target.py ) > gpurun_out/r2c14_memcheck_target.log 2>&1
grep -v "Host Frame" gpurun_out/r2c14_memcheck_target.log | head -8 | cut -c1-240
( time timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench_err.log
tail -c 600 gpurun_out/r2c14_bench.json
