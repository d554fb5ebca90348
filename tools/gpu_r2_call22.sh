#!/bin/bash
# round 2, GPU call 22 (1 GPU): fused store_solution + lean result copy -- GPU suite, smoke, bench line, SVM 1-GPU solve time
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c22_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2c22_pytest.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/r2c22_pytest.log | head
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c22_smoke.log 2>&1; tail -2 gpurun_out/r2c22_smoke.log
( time timeout 400 python bench.py --same-config-budget 30 ) > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c22_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["time_to_solution_ms"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"], d["ref_cuda"].get("b200_speedup_solve"))
PY
( timeout 200 python tools/svm_1gpu_probe.py 1.0 block ) > gpurun_out/r2c22_svm_block.log 2>&1; cut -c1-330 gpurun_out/r2c22_svm_block.log
