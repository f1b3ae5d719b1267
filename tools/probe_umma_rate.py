import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpu_tiled_qr_decomposition_b200 as q
reps = 256
out = np.zeros(8, dtype=np.int64)
L = q.probe_lib()
L.b200qr_probe_umma_rate.argtypes = [C.c_void_p, C.c_int]
for _ in range(2):
    q.check(L.b200qr_probe_umma_rate(out.ctypes.data_as(C.c_void_p), reps), "probe")
names = ["N=32 1acc", "N=32 2acc", "N=64 1acc", "N=64 2acc", "N=128 1acc", "N=128 2acc", "N=256 1acc", "-"]
for n, v in zip(names, out):
    print("%-12s %8d cycles  %.1f per MMA" % (n, v, v / (reps * 4.0)))
out2 = np.zeros(2, dtype=np.int64)
L.b200qr_probe_umma_rate2.argtypes = [C.c_void_p, C.c_int]
for _ in range(2):
    q.check(L.b200qr_probe_umma_rate2(out2.ctypes.data_as(C.c_void_p), reps), "probe2")
print("one issuing thread : %d cycles for %d MMAs (%.1f each)" % (out2[0], reps * 4, out2[0] / (reps * 4.0)))
print("two issuing threads: %d cycles for %d MMAs (%.1f cycles per MMA overall)" % (out2[1], reps * 8, out2[1] / (reps * 8.0)))
