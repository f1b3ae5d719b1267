"""Panel-kernel latency / throughput on one SM (libqrb200_probe.so, tools/qrb200_probe.h): SM cycles per 32x32 tile of the
one-warp GEQRT / TSQRT alone, with eight FFMA-bound warps on the SM, and with 2..16 independent chains sharing the SM.
Round-1 kernels measured 30.9k / 15.3k cycles (GEQRT / TSQRT alone), 21.2k per TSQRT with 8 chains, 33.5k with 16."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_tiled_qr_decomposition_b200 as q

P = q.probe_lib()
reps = 50
for nw, busy in ((1, 0), (1, 1), (2, 0), (4, 0), (8, 0), (12, 0), (16, 0)):
    cyc = (C.c_longlong * 32)()
    out = np.zeros(3 * 1024, dtype=np.float32)
    rc = P.b200qr_probe_panel(nw, busy, reps, cyc, out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    g = [cyc[2 * w] for w in range(nw)]
    t = [cyc[2 * w + 1] for w in range(nw)]
    print("%2d chains%s: GEQRT %6d..%6d  TSQRT %6d..%6d cycles/tile   (%.1f TSQRT tiles per 100k SM cycles)"
          % (nw, " + 8 FFMA warps" if busy else "", min(g), max(g), min(t), max(t), sum(1e5 / c for c in t)))
    if nw == 1 and not busy:
        # sanity: the GEQRT tile is a QR of the input up to rounding: |R| column norms = input column norms
        a = out[:1024].reshape(32, 32)
        print("   GEQRT R diagonal (first 4): %s" % np.array2string(np.diag(a)[:4], precision=4))
