"""DRAM access-pattern probe (libqrb200_probe.so): a read + write pass over n x n float32 in the tensor-core engine's sweep
order, column-major (a tile = 32 lines ld*4 bytes apart) against tile-packed (a tile = 4 KB contiguous)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_tiled_qr_decomposition_b200 as q

P = q.probe_lib()
for n in (16384, 32768):
    for packed in (0, 1):
        for cps in (2, 4, 8):
            ms = C.c_float(0)
            rc = P.b200qr_probe_dram_pattern(n, packed, cps, C.byref(ms))
            assert rc == 0, rc
            print("n %5d  %-12s %d CTAs/SM: %7.3f ms  %.2f TB/s (read + write)" % (
                n, "tile-packed" if packed else "column-major", cps, ms.value, 8.0 * n * n / ms.value * 1e-9))
