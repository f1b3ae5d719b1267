#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/s33
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -5 ${T}_tests.log
