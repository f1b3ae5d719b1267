set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s10_gpus.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29701 tools/dist_scan.py 32768 "-:-" "48:-" "128:-" "-:8" > gpurun_out/s10_scan8.log 2>&1
timeout 400 $TR --master-port 29702 tools/dist_stress.py 8192 100 > gpurun_out/s10_stress8.log 2>&1
timeout 900 $TR --master-port 29703 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/s10_bench8.log 2> gpurun_out/s10_bench8.err
echo "bench rc $?" >> gpurun_out/s10_bench8.err
grep -h "world\|stress" gpurun_out/s10_scan8.log gpurun_out/s10_stress8.log
