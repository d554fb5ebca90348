#!/bin/bash
# round 2, GPU call 18 (2 GPUs): the sharded parity suite on the final libraries
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 ) > gpurun_out/r2c18_pytest_sharded.log 2>&1
tail -6 gpurun_out/r2c18_pytest_sharded.log; grep -E "^FAILED|^E  " gpurun_out/r2c18_pytest_sharded.log | head -8 | cut -c1-300
