#!/bin/bash
# A/B: next-tile flag check + prefetch folded into the TSQRT mid hook (new) vs at the loop top behind a fence (base).
mkdir -p gpurun_out
T=gpurun_out/s31
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_dist.py tests/test_gpu_overlap.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -2 ${T}_tests.log
for rep in 1 2; do
  for n in 8192 12288; do
    QRB200_TC=0 timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed 's/^/simt new  /' >> ${T}_ab.log
    QRB200_TC=0 QRB200_LIB=$PWD/oldlib/libqrb200_base.so timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed "s/^/simt base /" >> ${T}_ab.log
  done
  for n in 8192 16384 32768; do
    timeout 300 python tools/tc3_time.py $n 3 2>&1 | sed 's/^/tc3  new  /' >> ${T}_ab.log
    QRB200_LIB=$PWD/oldlib/libqrb200_base.so timeout 300 python tools/tc3_time.py $n 3 2>&1 | sed "s/^/tc3  base /" >> ${T}_ab.log
  done
done
cat ${T}_ab.log
