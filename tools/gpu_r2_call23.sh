#!/bin/bash
# round 2, GPU call 23 (1 GPU): fused store_solution restricted to the solver's own solution arrays -- GPU suite, smoke
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c23_pytest.log 2>&1
grep -E "passed|failed|Segmentation" gpurun_out/r2c23_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2c23_pytest.log | head
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c23_smoke.log 2>&1; tail -2 gpurun_out/r2c23_smoke.log
