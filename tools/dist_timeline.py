"""Per-rank timeline of the multi-GPU factorisation (torchrun, trace mode 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rk = q.DistRank(n, n, rank, world, trace=2)
blobs = [None] * world
dist.all_gather_object(blobs, rk.export())
rk.connect(blobs)
g = torch.Generator(device="cuda").manual_seed(5)
A0 = torch.rand((n, n), generator=g, device="cuda") * 2 - 1
s = torch.cuda.current_stream().cuda_stream
for it in range(2):
    rk.upload(A0.data_ptr()); rk.reset(s); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rk.factor(s); e1.record(); torch.cuda.synchronize(); dist.barrier()
ms = e0.elapsed_time(e1)
tr = rk.read_trace().astype(np.int64)
p = n // 32
G = rk.block
mine = tr[tr[:, 4] >= 0] if False else tr
# per step blocks
off = 0
lines = []
t0 = None
tot_wait = 0.0; tot_rows = 0; waits_by_dist = {}
pan_start = {}
for k in range(p):
    cnt = (p - k) * (p - k)
    blk = tr[off:off + cnt].reshape(p - k, p - k, 6)
    off += cnt
    st, fi = blk[:, :, 4], blk[:, :, 5]
    if (k // G) % world == rank:
        if t0 is None: t0 = st[0, 0]
        pan_start[k] = (st[0, 0], fi[0, -1])
    for jj in range(1, p - k):
        j = k + jj
        if (j // G) % world != rank: continue
        if j % G != 0 and jj != 1: continue          # one column per group is enough (first of group, or k+1)
        s_ = st[jj, 1:].astype(np.float64)
        if len(s_) < 3: continue
        gaps = np.diff(s_) / 1e3
        d = (j // G) - ((k + 1) // G)
        w = np.clip(gaps - 9.5, 0, None).sum()
        waits_by_dist.setdefault(min(d, 4), [0.0, 0.0])
        waits_by_dist[min(d, 4)][0] += w
        waits_by_dist[min(d, 4)][1] += gaps.sum()
res = {"rank": rank, "ms": ms, "npanel_ctas?": None, "grid": rk.grid,
       "wait_frac_by_group_distance": {d: round(v[0] / max(v[1], 1e-9), 3) for d, v in sorted(waits_by_dist.items())},
       "sweep_time_share_by_distance": {d: round(v[1], 0) for d, v in sorted(waits_by_dist.items())}}
ks = sorted(pan_start)
if ks:
    base = pan_start[ks[0]][0]
    res["panels"] = [(k, round((pan_start[k][0] - base) / 1e3), round((pan_start[k][1] - base) / 1e3)) for k in ks[:: max(1, len(ks) // 12)]]
out = [None] * world
dist.all_gather_object(out, res)
if rank == 0:
    for r in out: print(r)
dist.destroy_process_group()
