"""tcgen05 building blocks (csrc/qr_umma.cuh, csrc/probes/umma_probe.cu): TMA 128B-swizzled tiles ->
3xTF32 split -> tcgen05.mma kind::tf32 into TMEM -> tcgen05.ld, against float64 references.
Stated tolerance: 3xTF32 keeps ~2^-21 per product; |err| <= 2e-6 * K * max|a| * max|b|."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_umma_products_match_float64(cuda, qr):
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, (128, 32)).astype(np.float32)        # storage (col, row): 32 x 128 column-major
    V = rng.uniform(-1, 1, (32, 32)).astype(np.float32)         # storage (col, row)
    W = rng.uniform(-1, 1, (128, 32)).astype(np.float32)        # row-major 128 x 32
    Z = np.zeros((128, 32), dtype=np.float32)
    D2 = np.zeros((128, 32), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    qr.check(qr.probe_lib().b200qr_probe_umma(p(A), p(V), p(W), p(Z), p(D2)), "b200qr_probe_umma")
    Am = A.T.astype(np.float64)                                  # (32 rows, 128 cols)
    Vm = V.T.astype(np.float64)                                  # (32 rows, 32 reflectors)
    Zref = Am.T @ Vm                                             # Z[m][n] = sum_r A[r][m] V[r][n]
    assert np.max(np.abs(Z - Zref)) <= 2e-6 * 32
    D2ref = Am.T - W.astype(np.float64) @ Vm.T                   # A^T - W V^T
    assert np.max(np.abs(D2 - D2ref)) <= 2e-6 * 32
    # and it is far better than a single TF32 pass would be (~1e-3 * 32 * 0.3)
    assert np.max(np.abs(Z - Zref)) < 1e-4
