set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/s19_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s19_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s19_bench.log 2> gpurun_out/s19_bench.err
echo "bench rc $?" >> gpurun_out/s19_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s19_bench_ref.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1500 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/s19_launches_bench.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1024 > gpurun_out/s19_ncu_bench.out 2>&1
tail -3 gpurun_out/s19_pytest.log; tail -2 gpurun_out/s19_smoke.log
