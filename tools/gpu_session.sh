set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "trace" > gpurun_out/s18_pytest.log 2>&1
for n in 8192 32768; do
  python tools/tc3_time.py $n 3 > gpurun_out/s18_new_${n}.log 2>&1
  QRB200_LIB=$PWD/oldlib/libqrb200_head.so python tools/tc3_time.py $n 3 > gpurun_out/s18_head_${n}.log 2>&1
done
tail -2 gpurun_out/s18_pytest.log
for f in gpurun_out/s18_*_*.log; do echo "$f: $(tail -1 $f)"; done
