#!/bin/bash
# round 2, GPU call 20 (1 GPU): SVM single-GPU solve time, block-seeded vs global generator
mkdir -p gpurun_out
( time timeout 420 python tools/svm_1gpu_probe.py 1.0 ) > gpurun_out/r2c20_svm_probe.log 2>&1
cat gpurun_out/r2c20_svm_probe.log | cut -c1-700
