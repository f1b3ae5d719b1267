"""Side benchmarks for BASELINE.json configs[2] (batched 10 000 x 256x256) and configs[3] (TSQR
4 194 304 x 256), single GPU; prints one JSON line each.  (bench.py keeps the headline contract.)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_tiled_qr_decomposition_b200 as q

L = q.lib()
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, restore, reps=3, warm=2):
    for _ in range(warm):
        restore(); fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


def batched(nprob=10000, m=256, n=256):
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((nprob, n, m), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    fl = C.c_longlong(0)
    q.check(L.b200qr_batched_workspace_floats(m, n, nprob, C.byref(fl)))
    tau = torch.zeros(fl.value, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    fn = lambda: q.check(L.b200qr_batched_factor(C.c_void_p(d.data_ptr()), m, n, m, m * n, nprob, C.c_void_p(tau.data_ptr()),
                                                 C.c_void_p(ctr.data_ptr()), C.c_void_p(stream)))
    ms = timeit(fn, lambda: d.copy_(A0))
    flops = nprob * (2.0 * m * n * n - 2.0 * n ** 3 / 3)
    # check one problem: R^T R = A^T A
    b = nprob // 2
    R = torch.triu(d[b].t().double()); Ad = A0[b].t().double()
    err = (torch.linalg.norm(R.t() @ R - Ad.t() @ Ad) / torch.linalg.norm(Ad.t() @ Ad)).item()
    print(json.dumps({"workload": "batched %d x (%dx%d), b=32" % (nprob, m, n), "ms": ms, "gflops": flops / ms * 1e-6,
                      "hbm_algorithmic_gbs": 8.0 * nprob * m * n / ms * 1e-6, "gram_rel_err": err}))


def tsqr(m=4194304, n=256, leaf=2048, fanin=4):
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    h = C.c_void_p()
    q.check(L.b200qr_tsqr_create(C.byref(h), m, n, leaf, fanin))
    R = torch.zeros((n, n), device="cuda")
    fn = lambda: q.check(L.b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()), C.c_void_p(stream)))
    ms = timeit(fn, lambda: d.copy_(A0))
    flops = 2.0 * m * n * n - 2.0 * n ** 3 / 3
    Rm = R.t().double()
    G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for c in range(0, m, 1 << 20):
        blk = A0[:, c:c + (1 << 20)].double()
        G += blk @ blk.t()
    err = (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item()
    print(json.dumps({"workload": "TSQR %dx%d, leaf %d rows, fan-in %d, %d tree levels" % (m, n, leaf, fanin, L.b200qr_tsqr_levels(h)),
                      "ms": ms, "gflops": flops / ms * 1e-6, "hbm_algorithmic_gbs": 8.0 * m * n / ms * 1e-6, "gram_rel_err": err}))
    L.b200qr_tsqr_destroy(h)


def tsqr_streaming(m=4194304, n=256, strips=0):
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    h = C.c_void_p()
    q.check(L.b200qr_tsqr_create_streaming(C.byref(h), m, n, strips))
    R = torch.zeros((n, n), device="cuda")
    fn = lambda: q.check(L.b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()), C.c_void_p(stream)))
    ms = timeit(fn, lambda: d.copy_(A0))
    flops = 2.0 * m * n * n - 2.0 * n ** 3 / 3
    Rm = R.t().double()
    G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for c in range(0, m, 1 << 20):
        blk = A0[:, c:c + (1 << 20)].double()
        G += blk @ blk.t()
    err = (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item()
    print(json.dumps({"workload": "TSQR %dx%d, streaming engine, %d strips, %d levels" % (m, n, strips, L.b200qr_tsqr_levels(h)),
                      "ms": ms, "gflops": flops / ms * 1e-6, "hbm_algorithmic_gbs": 8.0 * m * n / ms * 1e-6, "gram_rel_err": err}))
    L.b200qr_tsqr_destroy(h)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "batched"):
        batched()
    if which in ("all", "tsqr"):
        for leaf in (512, 2048):
            tsqr(leaf=leaf)
    if which in ("all", "tsqr", "stream"):
        for early, nb, rb in (("1", "1", "48"), ("1", "1", "32"), ("0", "1", "48"), ("0", "8", "48")):
            # early broadcast of the next column (two barriers) / reflectors per barrier / block rows
            os.environ["QRB200_TSQR_EARLY"], os.environ["QRB200_TSQR_NB"], os.environ["QRB200_TSQR_RB"] = early, nb, rb
            print("QRB200_TSQR_EARLY=%s QRB200_TSQR_NB=%s QRB200_TSQR_RB=%s" % (early, nb, rb), end="  ")
            tsqr_streaming(strips=0)
        for k in ("QRB200_TSQR_EARLY", "QRB200_TSQR_NB", "QRB200_TSQR_RB"):
            os.environ.pop(k, None)
        tsqr_streaming(m=1048576)
