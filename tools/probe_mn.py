"""Does tcgen05.mma kind::tf32 accept the reflector tile as an MN-major B operand?  (umma_probe mode bit 8)"""
import os, sys, ctypes as C
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
ref = W.astype(np.float64) @ Vm.T
for var in range(4):
    Z = np.zeros((128, 32), dtype=np.float32); D2 = np.zeros((128, 32), dtype=np.float32)
    mode = 1 | 2 | 8 | (var << 4)
    q.check(L.b200qr_probe_umma_mode(p(A), p(V), p(W), p(Z), p(D2), mode), "probe")
    print("variant", var, "max|P - W V^T| =", np.max(np.abs(D2 - ref)), " max|P| =", np.max(np.abs(D2)))
Z = np.zeros((128, 32), dtype=np.float32); D2 = np.zeros((128, 32), dtype=np.float32)
q.check(L.b200qr_probe_umma_mode(p(A), p(V), p(W), p(Z), p(D2), 3), "probe")
print("K-major V^T copy: max|P - W V^T| =", np.max(np.abs(D2 - ref)))
