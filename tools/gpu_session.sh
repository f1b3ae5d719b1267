set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/s3_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/s3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1
tail -3 gpurun_out/s3_pytest.log
