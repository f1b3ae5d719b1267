"""Time (CUDA events, best of 3) one streaming-TSQR factorisation: python tools/tsqr_time.py [m] [n] [strips].
Also the command the ncu capture of tsqr_strip_kernel is taken from (profiles/)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_tiled_qr_decomposition_b200 as q

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
strips = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(5)
A0 = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
d = torch.empty_like(A0)
R = torch.zeros((n, n), device="cuda")
t = q.Tsqr(m, n, strips=strips)
s = torch.cuda.current_stream().cuda_stream
best = 1e30
for it in range(4):
    d.copy_(A0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t.factor(d.data_ptr(), m, R.data_ptr(), s); e1.record()
    torch.cuda.synchronize()
    if it:
        best = min(best, e0.elapsed_time(e1))
fl = 2.0 * m * n * n - 2.0 * n ** 3 / 3
print("TSQR %d x %d streaming, %d levels: %.3f ms  %.2f TFLOP/s  %.1f GB/s algorithmic" %
      (m, n, t.levels, best, fl / best * 1e-9, 8.0 * m * n / best * 1e-6))
t.close()
