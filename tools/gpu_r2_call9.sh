#!/bin/bash
# round 2, GPU call 9 (1 GPU): final state -- GPU suite, smoke, compute-sanitizer, bench line, ncu evidence,
# the sharded arm's workloads on ONE GPU (N = 1 under torchrun env)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r2c9_pytest.log 2>&1
tail -4 gpurun_out/r2c9_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c9_smoke.log 2>&1
tail -6 gpurun_out/r2c9_smoke.log
# ---- compute-sanitizer on the final code
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_batch.py -q --timeout 550 -k "not larger_than_one_wave and not settings1 and not settings2 and not settings3" ) > gpurun_out/r2c9_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2c9_memcheck.log | tail -4
( time timeout 400 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_lin_alg.py tests/test_gpu_solve_parity.py -q --timeout 380 -k "submatrix or polish or golden_solutions or large_qp" ) > gpurun_out/r2c9_memcheck_polish.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2c9_memcheck_polish.log | tail -4
( time timeout 500 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_batch.py -q --timeout 450 -k "pcg or residual or determinism" ) > gpurun_out/r2c9_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2c9_racecheck.log | tail -4
# ---- the bench line (default flags)
( time timeout 800 python bench.py ) > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench_err.log
tail -c 1200 gpurun_out/r2c9_bench.json; tail -3 gpurun_out/r2c9_bench_err.log
# ---- ncu: launch list of the bench command (host loop: ncu cannot see kernel nodes of a conditional graph), full capture of the dominant pass and of the batch kernel
( B200_PCG_HOSTLOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/r2c9_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2c9_bench_under_ncu.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:g_lean_pass -s 8 -c 2 -f -o gpurun_out/r2c9_leanpass python tools/gpu_ncu_target.py 1.0 3 ) > gpurun_out/r2c9_ncu_leanpass.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_admm -c 1 -f -o gpurun_out/r2c9_batch python tools/batch_mpc.py 2048 --cpu-sample 0 ) > gpurun_out/r2c9_ncu_batch.log 2>&1
ls -la gpurun_out | grep r2c9
# ---- configs[3] on ONE GPU through the sharded arm (N = 1)
( time RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29571 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r2c9_bench_svm_1gpu.json 2> gpurun_out/r2c9_bench_svm_1gpu_err.log
( time RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29571 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 1 --workload huber ) > gpurun_out/r2c9_bench_huber_1gpu.json 2> gpurun_out/r2c9_bench_huber_1gpu_err.log
python - <<'PY'
import json
for f in ("r2c9_bench_svm_1gpu.json", "r2c9_bench_huber_1gpu.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["status"], d["obj_val"], d["admm_iters_per_step"], d["cg_iters_per_admm_iter"], d["e2e"]["time_to_solution_ms"], d.get("roofline", {}).get("phases_us"))
    except Exception as e:
        print(f, "FAILED", e)
PY
