#!/bin/bash
# Two GPUs after the panel-CTA model change: SIMT engine 8192^2 (default = new count, and the old one), tensor-core 32768^2,
# the multi-GPU tests.
mkdir -p gpurun_out
T=gpurun_out/s56
QRB200_TC=0 QRB200_WATCHDOG_MS=30000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/dist_scan.py 8192 "-:-" "-:14" "-:10" "-:5" > ${T}_simt.log 2>&1
grep "^world" ${T}_simt.log
QRB200_WATCHDOG_MS=30000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/dist_scan.py 32768 "-:-" > ${T}_tc.log 2>&1
grep "^world" ${T}_tc.log
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -2 ${T}_tests.log
