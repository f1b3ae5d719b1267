set -x
mkdir -p gpurun_out
for rep in 1 2; do
for n in 8192 16384 32768; do
  python tools/tc3_time.py $n 3 > gpurun_out/s21_new_${n}_$rep.log 2>&1
  QRB200_LIB=$PWD/oldlib/libqrb200_head.so python tools/tc3_time.py $n 3 > gpurun_out/s21_head_${n}_$rep.log 2>&1
done
done
for f in gpurun_out/s21_*_*.log; do echo "$f: $(tail -1 $f)"; done
