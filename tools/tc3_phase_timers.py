"""Phase timers of the tensor-core engine 3 (opts.trace == 3).  usage: python tools/tc3_phase_timers.py [n]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
os.environ["QRB200_TC"] = "3"
torch.cuda.set_device(0)
A = torch.rand((n, n), device="cuda") * 2 - 1
plan = q.Plan(n, n, trace=3)
for it in range(2):
    d = A.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream); e1.record()
    torch.cuda.synchronize()
out = np.zeros(32, dtype=np.uint64)
q.check(q.lib().b200qr_plan_debug_counters(plan._h, out.ctypes.data_as(C.c_void_p), 32))
v = [float(x) for x in out]
rows = v[4]
us = lambda x: x / rows / 1965.0   # SM cycles at 1965 MHz -> us
print("variant", plan.variant, "grid", plan.grid, "kernel ms %.2f" % e0.elapsed_time(e1), "rows (all CTAs)", rows)
print("per row (us): chain total %.2f | waits: YR %.2f WR %.2f PR(prev) %.2f -> chain compute %.2f" % (
    us(v[0]), us(v[1]), us(v[2]), us(v[3]), us(v[0] - v[1] - v[2] - v[3])))
print("  loader: total %.2f (non-blocking round robin over its four streams)" % us(v[8]))
print("  pre: CF %.2f PF %.2f | post: PR %.2f OUT-stage %.2f | storer: CC %.2f wait_group %.2f retire %.2f" % (us(v[9]), us(v[10]), us(v[11]), us(v[6]), us(v[12]), us(v[13]), us(v[20])))
print("  issuer A: VF %.2f PS %.2f | issuer B: SR %.2f WS %.2f" % (us(v[14]), us(v[15]), us(v[16]), us(v[17])))
print("  chain detail: ld W %.2f  split/st W %.2f  ld Y %.2f  split/st S %.2f  fence+arrive SR %.2f  arrive WS %.2f" % tuple(us(v[i]) for i in range(21, 27)))
print("  loader batches per row %.2f" % (v[27] / rows))
print("  issuer A: mma issue %.2f | issuer B: m2 issue %.2f  m3 issue %.2f" % tuple(us(v[i]) for i in range(29, 32)))
print("  fenced polls %.0f" % v[27])
