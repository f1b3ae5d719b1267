#!/bin/bash
# Eight GPUs: the bench line at N=8 (value, e2e through b200qr_dist_factor_host, TSQR over 8 GPUs).
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_session.sh'
mkdir -p gpurun_out
T=gpurun_out/s35
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 > ${T}_bench8.log 2>&1
grep '^{' ${T}_bench8.log | cut -c1-400; tail -3 ${T}_bench8.log | cut -c1-300
