set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/s20_ncu_batched.csv -k regex:batched python bench.py --only extras --extras batched --no-cpu --steps 1 > gpurun_out/s20_ncu_batched.out 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/s20_ncu_tsqr.csv -k regex:"tsqr_strip" python bench.py --only extras --extras tsqr --no-cpu --steps 1 > gpurun_out/s20_ncu_tsqr.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:batched_warp -c 1 -f -o gpurun_out/s20_batched python bench.py --only extras --extras batched --no-cpu --steps 1 > gpurun_out/s20_ncu_batched_full.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiledqr_persistent_tc3 -c 1 -f -o gpurun_out/s20_tc3_16384 python tools/tc3_time.py 16384 3 > gpurun_out/s20_ncu_tc3_full.out 2>&1
ls -la gpurun_out/*.ncu-rep
