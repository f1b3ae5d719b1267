import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpu_tiled_qr_decomposition_b200 as q
L = q.probe_lib()
L.b200qr_probe_umma_mode.argtypes = [C.c_void_p] * 5 + [C.c_int]
rng = np.random.default_rng(7)
A = rng.uniform(-1, 1, (128, 32)).astype(np.float32)
V = rng.uniform(-1, 1, (32, 32)).astype(np.float32)
W = rng.uniform(-1, 1, (128, 32)).astype(np.float32)
p = lambda a: a.ctypes.data_as(C.c_void_p)
Am, Vm = A.T.astype(np.float64), V.T.astype(np.float64)
for mode, name in ((3, "A from smem"), (7, "A from TMEM")):
    Z = np.zeros((128, 32), np.float32); P = np.zeros((128, 32), np.float32)
    rc = L.b200qr_probe_umma_mode(p(A), p(V), p(W), p(Z), p(P), mode)
    print(name, "rc", rc, "Z err", np.abs(Z - Am.T @ Vm).max(), "P err", np.abs(P - W @ Vm.T).max())
