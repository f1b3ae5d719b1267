set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/s15_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s15_pytest.log
for n in 8192 16384 32768; do python tools/tc3_time.py $n 3 > gpurun_out/s15_tc3_$n.log 2>&1; done
python tools/tc3_time.py 8192 0 > gpurun_out/s15_simt_8192.log 2>&1
python tools/trace_timeline.py 8192 2 > gpurun_out/s15_trace_8192.log 2>&1
tail -3 gpurun_out/s15_pytest.log
for f in gpurun_out/s15_tc3_*.log gpurun_out/s15_simt_8192.log; do echo "$f: $(tail -1 $f)"; done; grep "lag\|GEQRT" gpurun_out/s15_trace_8192.log
