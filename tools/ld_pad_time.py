"""Does the leading dimension matter?  Factor n x n with ld = n + pad (pad in floats).  usage: python tools/ld_pad_time.py n pad [pad ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_tiled_qr_decomposition_b200 as q
n = int(sys.argv[1])
torch.cuda.set_device(0)
s = torch.cuda.current_stream().cuda_stream
plan = q.Plan(n, n)
for pad in [int(x) for x in sys.argv[2:]]:
    ld = n + pad
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.zeros((n, ld), device="cuda")
    A[:, :n] = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    best = 1e9
    for _ in range(3):
        d = A.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.factor(d.data_ptr(), ld, s); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("n %d variant %d ld = n + %d: %.2f ms  %.1f TFLOP/s" % (n, plan.variant, pad, best, (4 / 3) * n ** 3 / best / 1e9), flush=True)
    del A, d
