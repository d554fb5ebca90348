#!/bin/bash
# round 2, GPU call 19 (1 GPU): BASELINE configs on one B200 with the final libraries (Portfolio: see r02_portfolio_maxiter.md)
mkdir -p gpurun_out
( time timeout 540 python tools/config_table.py "Random" "Lasso" "MPC" "SVM" "Huber" ) > gpurun_out/r2c19_configs_table.txt 2> gpurun_out/r2c19_err.log
tail -8 gpurun_out/r2c19_configs_table.txt; tail -4 gpurun_out/r2c19_err.log
