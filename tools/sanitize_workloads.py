"""Small instances of every engine, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_workloads.py simt|tc|batched|tsqr|frontend
Each prints its own check against a float64 reference so a sanitizer run is also a correctness run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gpu_tiled_qr_decomposition_b200 as q

torch.cuda.set_device(0)
s = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "simt"
g = torch.Generator(device="cuda").manual_seed(1)


def gram(d, A, n):
    R = torch.triu(d.t()[:n].double())
    Ad = A.t().double()
    G = Ad.t() @ Ad
    return (torch.linalg.norm(R.t() @ R - G) / torch.linalg.norm(G)).item()


if which in ("simt", "tc"):
    m, n = (512, 384) if which == "simt" else (2048, 2048)
    A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    d = A.clone()
    plan = q.Plan(m, n, engine=1 if which == "simt" else 2, grid_ctas=0 if which == "tc" else 6)
    plan.factor(d.data_ptr(), m, s)
    torch.cuda.synchronize()
    print(which, "variant", plan.variant, "grid", plan.grid, "gram", gram(d, A, n))
elif which == "batched":
    import ctypes as C
    for nprob, env in ((40, "colwarp"), (700, "warp")):
        os.environ["QRB200_BATCHED"] = env
        m = n = 128
        A = torch.rand((nprob, n, m), generator=g, device="cuda") * 2 - 1
        d = A.clone()
        fl = C.c_longlong(0)
        q.check(q.lib().b200qr_batched_workspace_floats(m, n, nprob, C.byref(fl)))
        tau = torch.zeros(fl.value, device="cuda")
        ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
        q.check(q.lib().b200qr_batched_factor(C.c_void_p(d.data_ptr()), m, n, m, m * n, nprob, C.c_void_p(tau.data_ptr()),
                                              C.c_void_p(ctr.data_ptr()), C.c_void_p(s)))
        torch.cuda.synchronize()
        print("batched", env, nprob, "gram", max(gram(d[b], A[b], n) for b in (0, nprob - 1)))
elif which == "tsqr":
    for mode in ("stream", "tiles"):
        m, n = 32768, 128
        A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
        d = A.clone()
        t = q.Tsqr(m, n, strips=8, mode=mode)
        R = torch.zeros((n, n), device="cuda")
        t.factor(d.data_ptr(), m, R.data_ptr(), s)
        torch.cuda.synchronize()
        Rm, Ad = R.t().double(), A.t().double()
        G = Ad.t() @ Ad
        print("tsqr", mode, "gram", (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item())
        t.close()
elif which == "frontend":
    m, n = 300, 170
    A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    h = q.AnyQR(m, n)
    h.factor(A.data_ptr(), m, s)
    B = torch.rand((3, m), generator=g, device="cuda")
    X = torch.zeros((3, n), device="cuda")
    h.solve_ls(B.data_ptr(), m, 3, X.data_ptr(), n, s)
    torch.cuda.synchronize()
    want = torch.linalg.lstsq(A.t().double(), B.t().double()).solution
    print("frontend lstsq rel err", (torch.linalg.norm(X.t().double() - want) / torch.linalg.norm(want)).item())
    tp = q.TilePlan(256, 256, 64)
    A2 = torch.rand((256, 256), generator=g, device="cuda") * 2 - 1
    d2 = A2.clone()
    tp.factor(d2.data_ptr(), 256, s)
    torch.cuda.synchronize()
    print("tiles b=64 gram", gram(d2, A2, 256))
