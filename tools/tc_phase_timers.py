import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
os.environ["QRB200_TC"] = "2"
torch.cuda.set_device(0)
A = torch.rand((n, n), device="cuda") * 2 - 1
plan = q.Plan(n, n, trace=3)
for _ in range(2):
    d = A.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream); e1.record()
    torch.cuda.synchronize()
out = np.zeros(8, dtype=np.uint64)
q.check(q.lib().b200qr_plan_debug_counters(plan._h, out.ctypes.data_as(C.c_void_p), 8))
full, z, g, p_, tot, rows = [float(x) for x in out[:6]]
print("variant", plan.variant, "kernel ms", e0.elapsed_time(e1), "engines*rows", rows)
print("per row (us): total %.2f | wait FULL %.2f  ZREADY %.2f  GREADY %.2f  PREADY %.2f  -> other (compute, stores) %.2f" % (
    tot / rows / 1e3, full / rows / 1e3, z / rows / 1e3, g / rows / 1e3, p_ / rows / 1e3, (tot - full - z - g - p_) / rows / 1e3))
