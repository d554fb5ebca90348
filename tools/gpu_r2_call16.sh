#!/bin/bash
# round 2, GPU call 16 (1 GPU): final libraries -- GPU suite, smoke, batch, bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c16_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2c16_pytest.log | tail -2; grep -E "^FAILED" gpurun_out/r2c16_pytest.log | head
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c16_smoke.log 2>&1; tail -3 gpurun_out/r2c16_smoke.log
( timeout 200 python tools/batch_mpc.py 4096 ) 2>&1 | tail -1 | cut -c1-900 > gpurun_out/r2c16_batch_mpc.log; cat gpurun_out/r2c16_batch_mpc.log
( time timeout 600 python bench.py --same-config-budget 90 ) > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c16_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["time_to_solution_ms"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"], d["ref_cuda"].get("b200_speedup_solve"))
PY
