#!/bin/bash
# round 2, GPU call 21 (1 GPU): SVM single GPU, each generator in its own fresh process, launch tracer on
mkdir -p gpurun_out
( B200_TRACE_FILE=gpurun_out/r2c21_trace_global.csv timeout 300 python tools/svm_1gpu_probe.py 1.0 global ) > gpurun_out/r2c21_global.log 2>&1
( B200_TRACE_FILE=gpurun_out/r2c21_trace_block.csv timeout 200 python tools/svm_1gpu_probe.py 1.0 block ) > gpurun_out/r2c21_block.log 2>&1
cut -c1-420 gpurun_out/r2c21_global.log gpurun_out/r2c21_block.log
python tools/trace_by_tag.py gpurun_out/r2c21_trace_global.csv gpurun_out/r2c21_trace_block.csv > gpurun_out/r2c21_by_tag.txt 2>&1
cat gpurun_out/r2c21_by_tag.txt
