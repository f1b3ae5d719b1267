"""Repeat the single-GPU factorisation and compare every run with the first, bit for bit (a stage handed back too early,
a tile loaded before its producer's store landed or a progress word seen early would change the result).
usage: python tools/soak_single.py iters n[:engine] ...   (engine 0 auto, 1 SIMT, 2 tensor-core)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_tiled_qr_decomposition_b200 as q

iters = int(sys.argv[1])
torch.cuda.set_device(0)
s = torch.cuda.current_stream().cuda_stream
for spec in sys.argv[2:]:
    n, _, e = spec.partition(":")
    n, e = int(n), int(e or 0)
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    plan = q.Plan(n, n, engine=e)
    first, bad = None, 0
    for it in range(iters):
        d = A.clone()
        plan.factor(d.data_ptr(), n, s)
        torch.cuda.synchronize()
        if first is None:
            first = d
        elif not torch.equal(first, d):
            bad += 1
    print("n %d engine %d (variant %d): %d runs, %d differ from the first" % (n, e, plan.variant, iters, bad), flush=True)
    plan.close()
    del first, A
    torch.cuda.empty_cache()
