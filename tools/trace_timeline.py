"""Timeline analysis of the persistent scheduler (trace mode 2: %globaltimer per tile task)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.cuda.set_device(0)
A = torch.rand((n, n), device="cuda") * 2 - 1
plan = q.Plan(n, n, trace=2)
for _ in range(2):
    d = A.clone()
    plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
tr = plan.read_trace().astype(np.int64)
t0 = tr[:, 4].min()
st = {tuple(r[:4]): (int(r[4] - t0), int(r[5] - t0)) for r in tr}
p = n // 32
print("total span us", (tr[:, 5].max() - t0) / 1e3, "grid", plan.grid, "macro", plan.macro_tasks)
def chain(k, col, typ_first, typ_rest):
    out = []
    for i in range(k, p):
        ty = typ_first if i == k else typ_rest
        out.append(st[(ty, i, col, k)])
    return np.array(out)
for k in (0, 1, 2, 5, 10, 20, p // 2):
    if k >= p - 2: continue
    c = chain(k, k, 0, 2)
    dur = (c[:, 1] - c[:, 0])
    pace = np.diff(c[:, 1])
    print("P(%d): start %.1f us, end %.1f us, tile compute median %.2f us, pace median %.2f us p90 %.2f" % (
        k, c[0, 0] / 1e3, c[-1, 1] / 1e3, np.median(dur) / 1e3, np.median(pace) / 1e3, np.percentile(pace, 90) / 1e3))
    for col in (k + 1, min(p - 1, k + 20)):
        s = chain(k, col, 1, 3)
        dur = (s[1:, 1] - s[1:, 0]); pace = np.diff(s[1:, 1])
        if len(pace) == 0: continue
        lag = s[1:, 0] - c[1:, 1]
        print("   S col %d: start %.1f end %.1f, row start->publish median %.2f us, pace median %.2f p90 %.2f, lag behind panel finish median %.2f us" % (
            col, s[0, 0] / 1e3, s[-1, 1] / 1e3, np.median(dur) / 1e3, np.median(pace) / 1e3, np.percentile(pace, 90) / 1e3, np.median(lag) / 1e3))
