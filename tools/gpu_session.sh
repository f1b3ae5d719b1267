#!/bin/bash
# Round-end validation session: the GPU suite, smoke(), the default bench line and the reference arm.
# Run on a B200 box:  gpurun --timeout 3600 -- 'bash tools/gpu_session.sh'
mkdir -p gpurun_out
T=gpurun_out/s23
timeout 2400 python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > ${T}_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${T}_ref.log 2>&1
QRB200_CACHE=0 timeout 300 python -c "
import gpu_tiled_qr_decomposition_b200 as q, numpy as np, ctypes, time
lib=q.lib(); n=4096
a=np.asfortranarray(np.random.default_rng(0).standard_normal((n,n),dtype=np.float32))
for i in range(3):
    b=a.copy(order='F'); t=time.time()
    lib.cudaQRTask(b.ctypes.data_as(ctypes.c_void_p), n, n, 0, 0); print('call',i,round(time.time()-t,3), float(abs(b[0,0])))
" > ${T}_cache.log 2>&1
tail -3 ${T}_tests.log; cat ${T}_smoke.log | tail -2; tail -c 600 ${T}_bench.log
