#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/s49
timeout 120 python tools/stall_demo.py > ${T}_stall.log 2>&1; grep -v Warn ${T}_stall.log | tail -3
timeout 2400 python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -3 ${T}_tests.log
