"""SM cycles per 32x32 TSQRT tile when 1 / 2 / 4 / 8 / 12 / 16 independent one-warp chains share an SM
(b200qr_probe_panel_chains): the assumption behind the tall-skinny plan of DESIGN.md section 8."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_tiled_qr_decomposition_b200 as q

L = q.lib()
for nw in (1, 2, 4, 8, 12, 16):
    out = (C.c_longlong * 16)()
    q.check(L.b200qr_probe_panel_chains(out, 200, nw), "b200qr_probe_panel_chains")
    cyc = [out[i] for i in range(nw)]
    print("%2d chains: %6d .. %6d cycles per tile, %.1f tiles per 100k SM cycles" % (nw, min(cyc), max(cyc), sum(1e5 / c for c in cyc)))
