set -x
mkdir -p gpurun_out
for rb in 32 16; do
  QRB200_TSQR_RB=$rb python bench.py --only extras --extras tsqr --no-cpu --steps 5 > gpurun_out/s9_tsqr_v4_rb$rb.log 2>&1
  QRB200_TSQR_NOPERSIST=1 QRB200_TSQR_RB=$rb python bench.py --only extras --extras tsqr --no-cpu --steps 5 > gpurun_out/s9_tsqr_v4_rb${rb}_nopersist.log 2>&1
done
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/s9_ncu_tsqr.csv -k regex:"tsqr_strip" python bench.py --only extras --extras tsqr --no-cpu --steps 1 > gpurun_out/s9_ncu_tsqr.out 2>&1
timeout 600 python -m pytest tests/test_gpu_batched.py -m gpu -q -k tsqr > gpurun_out/s9_pytest.log 2>&1
