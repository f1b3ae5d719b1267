"""SIMT update inner loop on one SM (libqrb200_probe.so): SM cycles per tile row of 16 tiles with 1 / 2 / 4 matrix columns
per thread.  FFMA bound: 16 tiles x 2048 warp-FFMA / 4 schedulers = 8192 cycles."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_tiled_qr_decomposition_b200 as q

P = q.probe_lib()
for nc in (1, 2, 4):
    cyc = C.c_longlong(0)
    rc = P.b200qr_probe_sweep(nc, 200, C.byref(cyc))
    assert rc == 0, rc
    print("nc %d (%2d warps): %6d cycles per tile row of 16 tiles = %.3f us at 1.965 GHz, %.0f %% of the FFMA bound"
          % (nc, 16 // nc, cyc.value, cyc.value / 1965.0, 8192.0 / cyc.value * 100))
