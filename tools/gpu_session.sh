set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py -m gpu -q --maxfail=12 > gpurun_out/s5_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s5_pytest.log
python bench.py --only extras --extras sweep --no-cpu --steps 3 > gpurun_out/s5_sweep.log 2>&1
python tools/tc3_phase_timers.py 16384 > gpurun_out/s5_tc3_timers_16384.log 2>&1
QRB200_NC=1 python tools/trace_timeline.py 8192 > gpurun_out/s5_trace_8192_nc1.log 2>&1
QRB200_NC=1 python tools/tc3_time.py 8192 0 > gpurun_out/s5_time_8192_nc1.log 2>&1
python tools/tc3_time.py 8192 0 > gpurun_out/s5_time_8192_nc2.log 2>&1
python tools/tc3_time.py 8192 3 > gpurun_out/s5_time_8192_tc.log 2>&1
python tools/tc3_time.py 16384 3 > gpurun_out/s5_time_16384_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batched_warp -c 1 -f -o gpurun_out/s5_batched python bench.py --only extras --extras batched --no-cpu --steps 1 > gpurun_out/s5_ncu_batched.out 2>&1
export QRB200_WATCHDOG_MS=900000
for w in simt batched tsqr frontend; do
  for tool in memcheck racecheck synccheck; do
    timeout 420 compute-sanitizer --tool $tool python tools/sanitize_workloads.py $w > gpurun_out/s5_san_${tool}_$w.log 2>&1
    echo "rc $?" >> gpurun_out/s5_san_${tool}_$w.log
  done
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_workloads.py tc > gpurun_out/s5_san_memcheck_tc.log 2>&1
echo "rc $?" >> gpurun_out/s5_san_memcheck_tc.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_workloads.py tc > gpurun_out/s5_san_racecheck_tc.log 2>&1
echo "rc $?" >> gpurun_out/s5_san_racecheck_tc.log
tail -3 gpurun_out/s5_pytest.log
