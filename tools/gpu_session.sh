set -x
mkdir -p gpurun_out
for n in 16384 32768; do
  python tools/tc3_time.py $n 3 > gpurun_out/s11_tc3_new_$n.log 2>&1
  QRB200_LIB=$PWD/oldlib/libqrb200_head.so python tools/tc3_time.py $n 3 > gpurun_out/s11_tc3_head_$n.log 2>&1
done
python tools/tc3_time.py 8192 3 > gpurun_out/s11_tc3_new_8192.log 2>&1
QRB200_LIB=$PWD/oldlib/libqrb200_head.so python tools/tc3_time.py 8192 3 > gpurun_out/s11_tc3_head_8192.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/s11_pytest.log 2>&1
cat gpurun_out/s11_tc3_*.log
