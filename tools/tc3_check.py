"""Quick B200 check of the tensor-core engine 3: compare with the SIMT engine on the same input, time both.
usage: python tools/tc3_check.py [n] [tc-variant]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tcv = sys.argv[2] if len(sys.argv) > 2 else "3"
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(5)
A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
s = torch.cuda.current_stream().cuda_stream

def run(env):
    if env is None:
        os.environ.pop("QRB200_TC", None)
    else:
        os.environ["QRB200_TC"] = env
    plan = q.Plan(n, n)
    best = 1e9
    for _ in range(3):
        d = A.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.factor(d.data_ptr(), n, s); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    v = plan.variant
    # quality of the factorisation itself (independent of the other engine): Gram identity and Q applied by the GPU
    Ad = A.t().double()                                    # m x n
    R = torch.triu(d.t()).double()
    x = torch.randn((n, 8), device="cuda", dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(1))
    lhs, rhs = R.t() @ (R @ x), Ad.t() @ (Ad @ x)
    gram = (torch.linalg.norm(lhs - rhs) / torch.linalg.norm(rhs)).item()
    eps = float(np.finfo(np.float32).eps)
    X = torch.randn((32, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    Y = X.clone()
    plan.apply_q(d.data_ptr(), n, Y.data_ptr(), n, 32, False, s)
    plan.apply_q(d.data_ptr(), n, Y.data_ptr(), n, 32, True, s)
    torch.cuda.synchronize()
    orth = (torch.linalg.norm(Y.double() - X.double()) / (torch.linalg.norm(X.double()) * n * eps)).item()
    Xn = torch.randn((32, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    RX = (R @ Xn.t().double()).float().t().contiguous()
    plan.apply_q(d.data_ptr(), n, RX.data_ptr(), n, 32, False, s)
    torch.cuda.synchronize()
    AX = Ad @ Xn.t().double()
    back = (torch.linalg.norm(RX.t().double() - AX) / (torch.linalg.norm(Ad) * torch.linalg.norm(Xn.double()) / (32 ** 0.5) * n * eps)).item()
    print("   quality: gram_rel %.3g  orth/(n eps) %.3g  backward/(n eps) %.3g" % (gram, orth, back), flush=True)
    del Ad, R
    plan.close()
    return d, best, v

ref, t_ref, v_ref = run(None)
print("SIMT variant %d: %.3f ms  %.1f TFLOP/s" % (v_ref, t_ref, (4 / 3) * n ** 3 / t_ref / 1e9), flush=True)
out, t, v = run(tcv)
scale = ref.abs().max().item()
err = (out - ref).abs().max().item()
print("TC variant %d: %.3f ms  %.1f TFLOP/s   max|diff vs SIMT| = %.3g (scale %.3g, rel %.3g)" % (
    v, t, (4 / 3) * n ** 3 / t / 1e9, err, scale, err / scale), flush=True)
Rr, Rt = torch.triu(ref.t()), torch.triu(out.t())
sg = torch.sign((Rr * Rt).sum(dim=1))
print("rows of R with flipped sign:", int((sg < 0).sum().item()), " max|R - sign-aligned R_simt| rel = %.3g" % (
    ((Rt - sg[:, None] * Rr).abs().max() / scale).item()), flush=True)
bad = torch.nonzero((out - ref).abs() > 1e-3 * scale)
if bad.numel():
    print("first bad entries (col, row):", bad[:8].tolist(), "count", bad.shape[0])
    cols = torch.unique(bad[:, 0] // 32)[:16].tolist(); rows = torch.unique(bad[:, 1] // 32)[:16].tolist()
    print("bad tile cols", cols, "bad tile rows", rows)
