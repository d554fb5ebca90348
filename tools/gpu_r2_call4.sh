#!/bin/bash
# round 2, GPU call 4 (2 GPUs): row-sharded solve over peer memory -- parity tests, then the sharded bench at
# reduced scale with and without P2P
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c4_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 -k "p2p or block" ) > gpurun_out/r2c4_pytest_sharded.log 2>&1
tail -15 gpurun_out/r2c4_pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 600 $TR bench.py --gpus 2 --scale 0.1 --steps 3 --warmup 2 ) > gpurun_out/r2c4_bench_p2p_s01.json 2> gpurun_out/r2c4_bench_p2p_s01_err.log
( time B200_DIST_NO_P2P=1 timeout 600 $TR bench.py --gpus 2 --scale 0.1 --steps 3 --warmup 2 --no-strong-baseline ) > gpurun_out/r2c4_bench_nccl_s01.json 2> gpurun_out/r2c4_bench_nccl_s01_err.log
tail -c 2500 gpurun_out/r2c4_bench_p2p_s01.json; tail -5 gpurun_out/r2c4_bench_p2p_s01_err.log
tail -c 1500 gpurun_out/r2c4_bench_nccl_s01.json; tail -5 gpurun_out/r2c4_bench_nccl_s01_err.log
