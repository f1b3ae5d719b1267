#!/bin/bash
# Where the tensor-core engine overtakes the SIMT engine after the late SIMT ring changes.
mkdir -p gpurun_out
T=gpurun_out/s39
for n in 10240 12288 14336 16384; do
  QRB200_TC=0 timeout 300 python tools/tc3_time.py $n 0 2>&1 | sed 's/^/simt /' >> ${T}_cross.log
  timeout 300 python tools/tc3_time.py $n 3 2>&1 | sed 's/^/tc3  /' >> ${T}_cross.log
done
cat ${T}_cross.log
