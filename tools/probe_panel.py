import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_tiled_qr_decomposition_b200 as q
L = q.lib()
c = (C.c_longlong * 4)()
q.check(L.b200qr_probe_panel(c, 50))
print("cycles/tile: GEQRT alone %d, TSQRT alone %d; with 8 FFMA warps: GEQRT %d, TSQRT %d" % tuple(c))
