#!/bin/bash
# Eight GPUs: the bench line at N=8 (value, e2e through b200qr_dist_factor_host).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_session.sh'
mkdir -p gpurun_out
T=gpurun_out/s57
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --extras none --no-cpu > ${T}_bench8.log 2>&1
grep '^{' ${T}_bench8.log | cut -c1-300; tail -2 ${T}_bench8.log | cut -c1-300
