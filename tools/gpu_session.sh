set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_batched.py tests/test_gpu_tc.py -m gpu -q --maxfail=12 > gpurun_out/s6_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s6_pytest.log
python bench.py --only extras --extras batched --no-cpu --steps 5 > gpurun_out/s6_batched.log 2>&1
QRB200_TSQR_MODE=tiles python bench.py --only extras --extras tsqr --no-cpu --steps 3 > gpurun_out/s6_tsqr_tiles.log 2>&1
QRB200_NC=1 python tools/tc3_time.py 8192 0 > gpurun_out/s6_time_8192_nc1.log 2>&1
QRB200_NC=1 python tools/trace_timeline.py 8192 > gpurun_out/s6_trace_8192_nc1.log 2>&1
for n in 16384 32768; do
  python tools/tc3_time.py $n 3 > gpurun_out/s6_tc3_new_$n.log 2>&1
  QRB200_LIB=$PWD/oldlib/libqrb200_head.so python tools/tc3_time.py $n 3 > gpurun_out/s6_tc3_head_$n.log 2>&1
done
tail -3 gpurun_out/s6_pytest.log
