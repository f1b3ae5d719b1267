#!/bin/bash
# A/B: SIMT sweep storer releases its stage after the store's read (new) vs after publication (base).
mkdir -p gpurun_out
T=gpurun_out/s29
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_overlap.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -2 ${T}_tests.log
for rep in 1 2; do
  for n in 8192 12288; do
    QRB200_TC=0 timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed 's/^/new  /' >> ${T}_ab.log
    QRB200_TC=0 QRB200_LIB=$PWD/oldlib/libqrb200_st.so timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed "s/^/stor /" >> ${T}_ab.log
    QRB200_TC=0 QRB200_LIB=$PWD/oldlib/libqrb200_base.so timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed "s/^/base /" >> ${T}_ab.log
  done
done
cat ${T}_ab.log
