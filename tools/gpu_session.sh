#!/bin/bash
# A/B: L2 promotion of the tensor maps (QRB200_L2PROMO = 128 default / 256 / 64 / 0).
mkdir -p gpurun_out
T=gpurun_out/s50
for rep in 1 2; do
  for pr in 128 256 64; do
    QRB200_L2PROMO=$pr timeout 300 python tools/tc3_time.py 16384 3 2>&1 | grep -v Warn | sed "s/^/promo $pr /" >> ${T}_promo.log
  done
done
for pr in 128 256; do
  QRB200_L2PROMO=$pr timeout 300 python tools/tc3_time.py 32768 3 2>&1 | grep -v Warn | sed "s/^/promo $pr /" >> ${T}_promo.log
  QRB200_L2PROMO=$pr QRB200_TC=0 timeout 300 python tools/tc3_time.py 8192 0 2>&1 | grep -v Warn | sed "s/^/promo $pr /" >> ${T}_promo.log
  QRB200_L2PROMO=$pr QRB200_TC=0 timeout 300 python tools/tc3_time.py 12288 0 2>&1 | grep -v Warn | sed "s/^/promo $pr /" >> ${T}_promo.log
done
cat ${T}_promo.log
