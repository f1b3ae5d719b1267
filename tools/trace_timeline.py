"""Timeline analysis of the persistent scheduler (trace mode 2: %globaltimer per tile task)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
engine = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # 0 auto, 1 SIMT, 2 tensor-core (panel tasks only are marked)
torch.cuda.set_device(0)
A = torch.rand((n, n), device="cuda") * 2 - 1
plan = q.Plan(n, n, trace=2, engine=engine)
for _ in range(2):
    d = A.clone()
    plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
tr = plan.read_trace().astype(np.int64)
tr[:, 4:6] &= 0xffffffff                                  # low 32 bits of %globaltimer; unmarked tasks read 0xffffffff
marked = tr[:, 4] != 0xffffffff
t0 = tr[marked, 4].min()
print("variant", plan.variant, "marked tasks", int(marked.sum()), "of", len(tr))
if plan.variant == 5:
    # tensor-core engine: only the panel tasks carry marks
    p = n // 32
    off = 0
    starts, ends, geq, paces = [], [], [], []
    for k in range(p):
        cnt = (p - k) * (p - k)
        col = tr[off:off + (p - k)]                        # j = k: rows k .. p-1
        off += cnt
        st = (col[:, 4] - t0) / 1e3; fi = (col[:, 5] - t0) / 1e3
        starts.append(st[0]); ends.append(fi[-1]); geq.append(fi[0] - st[0])
        if p - k > 3:
            paces.append(np.median(np.diff(fi[1:])))
    starts = np.array(starts); lag = np.diff(starts)
    print("n", n, "span us %.1f" % max(ends))
    print("lag between consecutive panel starts (us): median %.1f, p10 %.1f, p90 %.1f; sum %.0f us" % (
        np.median(lag), np.percentile(lag, 10), np.percentile(lag, 90), lag.sum()))
    print("GEQRT duration median %.2f us; TSQRT pace of a chain median %.2f us" % (np.median(geq), np.median(paces)))
    for k in sorted(set([0, 1, 2, 8, 32, p // 4, p // 2, 3 * p // 4]) & set(range(p - 1))):
        print("P(%3d): start %9.1f end %9.1f  lag to next %.1f" % (k, starts[k], ends[k], lag[k]))
    sys.exit(0)
p = n // 32
span = (tr[:, 5].max() - t0) / 1e3
print("n", n, "total span us %.1f" % span, "grid", plan.grid, "macro", plan.macro_tasks)
# index: per step k block of (p-k)*(q-k) records ordered j-major then i (see b200qr_plan_read_trace)
off = 0
S = np.zeros((p, p, p, 2), dtype=np.float32) if n <= 2048 else None
rows = {}
off = 0
pan, far, near = [], [], []
for k in range(p):
    cnt = (p - k) * (p - k)
    blk = tr[off:off + cnt].reshape(p - k, p - k, 6)      # [j-k][i-k]
    off += cnt
    st = (blk[:, :, 4] - t0) / 1e3
    fi = (blk[:, :, 5] - t0) / 1e3
    pan.append((k, st[0, 0], fi[0, -1], np.median(np.diff(fi[0])) if p - k > 2 else 0))
    if p - k > 40:
        j1 = 1
        jf = p - k - 1
        near.append((k, st[j1, 0], fi[j1, -1], np.median(np.diff(fi[j1, 1:]))))
        far.append((k, st[jf, 0], fi[jf, -1], np.median(np.diff(fi[jf, 1:]))))
pan = np.array(pan); near = np.array(near); far = np.array(far)
for k in sorted(set([0, 1, 2, 8, 16, 32, p // 4, p // 2, 3 * p // 4, p - 2]) & set(range(p))):
    r = pan[k]
    line = "P(%3d): start %8.1f end %8.1f pace %.2f us" % (k, r[1], r[2], r[3])
    m = near[near[:, 0] == k]
    if len(m):
        line += " | near sweep: start %8.1f end %8.1f pace %.2f" % tuple(m[0, 1:])
        f = far[far[:, 0] == k][0]
        line += " | far sweep: start %8.1f end %8.1f pace %.2f" % tuple(f[1:])
    print(line)
# concurrency of panels over time
ts = np.linspace(0, span, 21)[1:-1]
conc = [(int(((pan[:, 1] <= t) & (pan[:, 2] >= t)).sum())) for t in ts]
print("active panels at 5%..95% of the run:", conc)
print("far-sweep pace median over steps: %.2f us; near-sweep pace median: %.2f us; panel pace median %.2f us" % (
    np.median(far[:, 3]), np.median(near[:, 3]), np.median(pan[:p - 2, 3])))
lag = np.diff(pan[:, 1])
print("lag between consecutive panel starts (us): median %.1f, p10 %.1f, p90 %.1f; sum %.0f us of %.0f us total" % (
    np.median(lag), np.percentile(lag, 10), np.percentile(lag, 90), lag.sum(), span))
geq = []
off = 0
for k in range(p):
    cnt = (p - k) * (p - k)
    r = tr[off]                       # (k, k, k): GEQRT
    geq.append((r[5] - r[4]) / 1e3)
    off += cnt
print("GEQRT duration median %.2f us" % np.median(geq))
