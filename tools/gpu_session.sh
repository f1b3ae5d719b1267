set -x
mkdir -p gpurun_out
python tools/tc_shapes_check.py 2 > gpurun_out/s2_shapes_new.log 2>&1
QRB200_LIB=$PWD/oldlib/libqrb200_r1.so QRB200_TC=3 python tools/tc_shapes_check.py 2 > gpurun_out/s2_shapes_r1.log 2>&1
python tools/probe_panel.py > gpurun_out/s2_probe.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s2_pytest.log
python tools/trace_timeline.py 8192 > gpurun_out/s2_trace_8192.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s2_bench.log 2> gpurun_out/s2_bench.err
echo "bench rc $?" >> gpurun_out/s2_bench.err
tail -3 gpurun_out/s2_pytest.log
