import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(5)
A = torch.rand((n, n), generator=g, device="cuda") * 2 - 1
def run(tc):
    if tc: os.environ["QRB200_TC"] = str(tc)
    else: os.environ.pop("QRB200_TC", None)
    plan = q.Plan(n, n)
    d = A.clone()
    plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    v = plan.variant
    plan.close()
    return d.cpu().numpy(), v
ref, v0 = run(0)
out, v1 = run(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
print("variants", v0, v1)
p = n // 32
diff = np.abs(out - ref)
diff[np.isnan(out)] = 1e9
T = diff.reshape(p, 32, p, 32).max(axis=(1, 3))       # [tile col][tile row]
bad = np.argwhere(T > 1e-2)
print("bad tiles", len(bad), "of", p * p, "nan count", int(np.isnan(out).sum()))
if len(bad):
    # first bad by column then row
    order = sorted((int(c), int(r)) for c, r in bad)
    print("first bad (tile col, tile row):", order[:12])
    rows = sorted(set(r for c, r in order)); cols = sorted(set(c for c, r in order))
    print("bad rows range", rows[0], rows[-1], "bad cols range", cols[0], cols[-1])
    c, r = order[0]
    print("tile diff sample", T[c, r], "out", out[c*32:(c*32+2), r*32:(r*32+4)], "ref", ref[c*32:(c*32+2), r*32:(r*32+4)])
