mkdir -p gpurun_out
python tools/probe_dram.py > gpurun_out/s22_dram.log 2>&1
cat gpurun_out/s22_dram.log
