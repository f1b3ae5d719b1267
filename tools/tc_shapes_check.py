"""Tensor-core engine against the oracle over several shapes (tall, wide, square); usage: tc_shapes_check.py [engine]
(QRB200_LIB selects another build of the library; with a round-1 library pass engine via QRB200_TC=3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch
import gpu_tiled_qr_decomposition_b200 as q
import pyoracle as po

po.build()
engine = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.cuda.set_device(0)
s = torch.cuda.current_stream().cuda_stream
for m, n in ((2048, 2048), (3072, 2048), (4096, 2048), (4096, 3072), (2048, 2560), (4096, 4096)):
    a = po.randzo(m, n)
    ref_out, _ = po.tiled_qr(a, m, n)
    scale = float(np.max(np.abs(np.triu(ref_out.T))))
    try:
        plan = q.Plan(m, n, engine=engine)
    except TypeError:
        plan = q.Plan(m, n)
    for rep in range(2):
        d = torch.from_numpy(a).cuda()
        plan.factor(d.data_ptr(), m, s)
        torch.cuda.synchronize()
        out = d.cpu().numpy()
        err = np.abs(out - ref_out)
        bad = np.argwhere(err > 5e-4 * scale)
        print("%dx%d variant %d grid %d rep %d: max err %.3g (scale %.3g)  bad elements %d%s" % (
            m, n, plan.variant, plan.grid, rep, err.max(), scale, len(bad),
            "  first bad (col,row) %s tile (%d,%d)" % (bad[0], bad[0][1] // 32, bad[0][0] // 32) if len(bad) else ""))
    plan.close()
