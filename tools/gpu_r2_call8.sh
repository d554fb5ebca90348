#!/bin/bash
# round 2, GPU call 8 (8 GPUs): where does an ADMM iteration of the 8-rank SVM solve go?  launch trace per rank
# (peer-memory path), and the NCCL path beside it
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
( time B200_TRACE_FILE=gpurun_out/r2c8_trace_p2p timeout 400 $TR tools/sharded_worker.py --family svm --scale 1.0 --blocks --reps 4 ) > gpurun_out/r2c8_worker_p2p.log 2>&1
grep -E "SHARDED" gpurun_out/r2c8_worker_p2p.log | cut -c1-900
( time B200_DIST_NO_P2P=1 timeout 400 $TR tools/sharded_worker.py --family svm --scale 1.0 --blocks --reps 4 ) > gpurun_out/r2c8_worker_nccl.log 2>&1
grep -E "SHARDED" gpurun_out/r2c8_worker_nccl.log | cut -c1-900
for r in 0 3 7; do echo "== rank $r"; python tools/gpu_phase_profile.py --summarise gpurun_out/r2c8_trace_p2p.rank$r | head -30; done
rm -f gpurun_out/r2c8_trace_p2p.rank*.last
ls -la gpurun_out | grep r2c8
