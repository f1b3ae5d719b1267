#!/bin/bash
# Round-end validation session: the GPU suite, smoke(), the default bench line and the reference arm.
# Run on a B200 box:  gpurun --timeout 3600 -- 'bash tools/gpu_session.sh'
# (multi-GPU: gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
#  --master-port 29517 bench.py --gpus N', and tests/test_gpu_dist.py on a box with at least two GPUs)
mkdir -p gpurun_out
T=gpurun_out/final
timeout 2400 python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${T}_ref.log 2>&1
tail -3 ${T}_tests.log; tail -2 ${T}_smoke.log; tail -c 600 ${T}_bench.log
