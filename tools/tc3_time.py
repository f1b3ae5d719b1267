"""Time one engine variant: python tools/tc3_time.py n [tc]  (env QRB200_SKEW / QRB200_NP honoured)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_tiled_qr_decomposition_b200 as q
n = int(sys.argv[1]); tcv = sys.argv[2] if len(sys.argv) > 2 else "3"
if tcv != "0": os.environ["QRB200_TC"] = tcv
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(5)
A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
plan = q.Plan(n, n)
s = torch.cuda.current_stream().cuda_stream
best = 1e9
for _ in range(3):
    d = A.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.factor(d.data_ptr(), n, s); e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("n %d variant %d skew %s np %s: %.2f ms  %.1f TFLOP/s" % (n, plan.variant, os.environ.get("QRB200_SKEW", "-"), os.environ.get("QRB200_NP", "-"),
      best, (4 / 3) * n ** 3 / best / 1e9), flush=True)
