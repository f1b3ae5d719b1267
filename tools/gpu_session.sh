#!/bin/bash
# Two GPUs: the multi-GPU tests a one-GPU box skips (incl. the host path with overlapped copies) and the bench line at N=2.
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_session.sh'
mkdir -p gpurun_out
T=gpurun_out/s34
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -5 ${T}_tests.log
QRB200_WATCHDOG_MS=30000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_tc3_check.py 16384 > ${T}_check.log 2>&1; grep -v Warning ${T}_check.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > ${T}_bench2.log 2>&1
grep '^{' ${T}_bench2.log | cut -c1-600; tail -3 ${T}_bench2.log | cut -c1-300
