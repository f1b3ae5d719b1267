#!/bin/bash
# Two GPUs: stress of both engines' multi-rank paths (repeated factorisations, bit-compared).
mkdir -p gpurun_out
T=gpurun_out/s38
QRB200_WATCHDOG_MS=30000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_stress.py 8192 60 1 > ${T}_stress_simt.log 2>&1
QRB200_WATCHDOG_MS=30000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/dist_stress.py 8192 60 2 > ${T}_stress_tc.log 2>&1
grep -v Warn ${T}_stress_simt.log | tail -2; grep -v Warn ${T}_stress_tc.log | tail -2
