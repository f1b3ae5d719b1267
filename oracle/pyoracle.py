"""ctypes front-end for the test oracle (oracle/liboracle.so, oracle/_ref/libref.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports it.

Matrices are column-major float32: a NumPy C-order array of shape (n, ldm) *is* the
m x n matrix with leading dimension ldm (qrdecomp.c:18).  Helpers below take/return
Fortran-style views `A[i, j]` built on that storage.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

TYPE_NAMES = ("QRS", "SAPP", "QRD", "DAPP")   # gridscheduler.h:10


def build(force=False):
    """Compile oracle/liboracle*.so (+ oracle/_ref when /root/reference exists)."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so")) \
        or not os.path.exists(os.path.join(HERE, "liboracle_fast.so"))
    if os.path.isfile("/root/reference/qrdecomp.c") and not os.path.exists(os.path.join(HERE, "_ref", "libref.so")):
        need = True
    if need:
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)
    # the reference's own self-test binary, linked against the B200 library (needs /root/reference and libqrb200.so)
    qlib = os.path.join(os.path.dirname(HERE), "gpu_tiled_qr_decomposition_b200", "libqrb200.so")
    exe = os.path.join(HERE, "_ref", "ref_selftest")
    if os.path.isfile("/root/reference/qrdecomp.c") and os.path.exists(qlib) and (
            force or not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(qlib)):
        subprocess.run(["make", "-C", HERE, "selftest"], check=True, capture_output=True)


def _bind(lib):
    lib.oracle_srand.argtypes = [C.c_uint]
    lib.oracle_fill_randzo.argtypes = [_f32p, C.c_int, C.c_int, C.c_int]
    lib.oracle_geqrt.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
    lib.oracle_tsqrt.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    lib.oracle_tsqrt_lead.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    lib.oracle_unmqr.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    lib.oracle_tsmqr.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    for name in ("oracle_tiled_qr_sched", "oracle_tiled_qr_static"):
        getattr(lib, name).argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.oracle_tiled_qr_threads.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.oracle_sched_trace.argtypes = [C.c_int, C.c_int, _i32p, C.c_long]
    lib.oracle_sched_trace.restype = C.c_long
    lib.oracle_apply_qt_f64.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _f64p, C.c_int, C.c_int]
    lib.oracle_now.restype = C.c_double
    return lib


_cache = {}


def lib(fast=False):
    key = "fast" if fast else "exact"
    if key not in _cache:
        build()
        _cache[key] = _bind(C.CDLL(os.path.join(HERE, "liboracle_fast.so" if fast else "liboracle.so")))
    return _cache[key]


def ref_lib(fast=False):
    """The unmodified reference CPU path (None if oracle/_ref was never built)."""
    key = "ref_fast" if fast else "ref"
    if key not in _cache:
        path = os.path.join(HERE, "_ref", "libref_fast.so" if fast else "libref.so")
        if not os.path.exists(path):
            _cache[key] = None
        else:
            r = C.CDLL(path)
            # qrdecomp.h:20,22-28 ; gridscheduler.h:19-21
            r.taskQRP_threads.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            r.initMatrix.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int]
            r.checkEqual.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
            r.checkEqual.restype = C.c_int
            r.initScheduler.argtypes = [C.c_int, C.c_int]
            r.initScheduler.restype = C.c_void_p
            r.getNextTask.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            r.doneATask.argtypes = [C.c_void_p, C.c_int, C.c_int, RefTask]
            r.doATask.argtypes = [RefTask, _f32p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int]
            _cache[key] = r
    return _cache[key]


class RefTask(C.Structure):      # gridscheduler.h:13-17
    _fields_ = [("taskType", C.c_int), ("l", C.c_int), ("m", C.c_int), ("k", C.c_int),
                ("taskStatus", C.c_int)]


def randzo(m, n, ldm=None, seed=5):
    """Reference input: srand(seed) then initMatrix(RANDZO) (qrdecomp.c:81,89,1383)."""
    ldm = ldm or m
    a = np.zeros((n, ldm), dtype=np.float32)
    L = lib()
    L.oracle_srand(seed)
    L.oracle_fill_randzo(a, m, n, ldm)
    return a


def tiled_qr(a, m, n, b=32, ldm=None, order="static"):
    """Factor a copy of `a`; returns (out, tau), both (n, ldm) float32 storage."""
    ldm = ldm or m
    out = np.ascontiguousarray(a, dtype=np.float32).copy()
    tau = np.zeros_like(out)
    fn = lib().oracle_tiled_qr_sched if order == "sched" else lib().oracle_tiled_qr_static
    fn(out, tau, m, n, b, ldm)
    return out, tau


def tsqr_stream(a, m, n, block_rows):
    """Tall-skinny QR by ONE TSQRT chain whose tile is as wide as the matrix -- the recurrence the streaming engine of
    the product runs per strip: R starts as the empty n x n triangle and every block of `block_rows` rows is one call of
    the TSQRT restatement (oracle_tsqrt = qrdecomp.c:689-721 STSQRF + 1243-1272 calcvkDouble + 647-687 updateDoubleQ_WY,
    whose tile sizes are parameters) with ma = n, mb = block_rows.  While fewer than n rows have gone by, a block only
    determines the reflectors k < rows seen so far (oracle_tsqrt_lead: the columns behind them are zero up to rounding,
    and both implementations set them to zero rather than let the sign of that noise pick the sign of a row of R).  `a`: (n, m) float32 storage of the column-major m x n
    matrix.  Returns (V, tau, R): V (n, m) storage with every block replaced by its reflector block V2, tau
    (m / block_rows, n), R (n, n) storage, upper triangular ([column, row])."""
    assert m % block_rows == 0
    L = lib()
    V = np.ascontiguousarray(a, dtype=np.float32).copy()
    nblk = m // block_rows
    tau = np.zeros((nblk, n), dtype=np.float32)
    ld = n + block_rows
    buf = np.zeros((n, ld), dtype=np.float32)                # [R ; block], one leading dimension for both (oracle_tsqrt)
    work = np.zeros(ld + 8, dtype=np.float32)
    for bi in range(nblk):
        rows = slice(bi * block_rows, (bi + 1) * block_rows)
        buf[:, n:] = V[:, rows]
        t = np.zeros(n, dtype=np.float32)
        top = buf.reshape(-1)
        nref = min(n, (bi + 1) * block_rows)
        L.oracle_tsqrt_lead(top, top[n:], t, n, block_rows, nref, n, ld, work)
        V[:, rows] = buf[:, n:]
        tau[bi] = t
    return V, tau, np.ascontiguousarray(buf[:, :n])


def ref_tiled_qr(a, m, n, ldm=None, fast=False):
    """The reference's own taskQRP_threads (8 pthreads, b = 32)."""
    ldm = ldm or m
    r = ref_lib(fast)
    src = np.ascontiguousarray(a, dtype=np.float32)
    out = np.zeros_like(src)
    tau = np.zeros_like(src)
    # the reference prints "CPU: x ms" to stdout from C; leave it be
    r.taskQRP_threads(src, out, tau, m, n, 32, ldm, 1)
    return out, tau


def sched_trace(M, N):
    cap = sum((M - k) * (N - k) for k in range(min(M, N)))
    buf = np.zeros((cap, 4), dtype=np.int32)
    cnt = lib().oracle_sched_trace(M, N, buf, cap)
    assert cnt == cap, (cnt, cap)
    return buf


def ref_sched_trace(M, N):
    r = ref_lib()
    grid = r.initScheduler(M, N)
    t = RefTask()
    out = []
    while r.getNextTask(C.byref(t), grid, M, N) == 0:
        out.append((t.taskType, t.l, t.m, t.k))
        r.doneATask(grid, M, N, t)
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def apply_qt_f64(out, tau, m, n, b, c, ldm=None):
    """Q^T @ c in float64 by replaying the flat tree.
    `c` is an (m, nc) ordinary NumPy matrix; returns an (m, nc) matrix."""
    ldm = ldm or m
    cc = np.ascontiguousarray(np.asarray(c, dtype=np.float64).T).copy()   # (nc, m) == col-major m x nc
    lib().oracle_apply_qt_f64(np.ascontiguousarray(out), np.ascontiguousarray(tau),
                              m, n, b, ldm, cc, cc.shape[0], m)
    return cc.T.copy()


def tau_from_v(out, m, n, b, ldm=None):
    """Recompute the tau matrix from stored reflectors: tau = 2/(1+|v_ess|^2)
    (qrdecomp.c:498-504, 660-664).  cudaQRTask never returns tau (gpucalc.cu:1610,1675)."""
    ldm = ldm or m
    A = out[:, :m].T.astype(np.float64)          # (m, n) view of values
    tau = np.zeros_like(out)
    M, N = m // b, n // b
    for k in range(min(M, N)):
        for l in range(k, M):
            blk = A[l * b:(l + 1) * b, k * b:(k + 1) * b]
            v = np.tril(blk, -1) if l == k else blk
            t = 2.0 / (1.0 + (v * v).sum(axis=0))
            tau[k * b, l * b:(l + 1) * b] = t.astype(np.float32)
    return tau


def quality(a, out, tau, m, n, b=32, ldm=None):
    """(backward, orthogonality, R) with Q rebuilt in fp64 from (V, tau):
    |A-QR|_F/(|A|_F n eps32), |Q^T Q - I|_F/(n eps32)."""
    ldm = ldm or m
    eps = float(np.finfo(np.float32).eps)
    A = a[:, :m].T.astype(np.float64)
    QtI = apply_qt_f64(out, tau, m, n, b, np.eye(m), ldm)      # Q^T
    R = np.triu(out[:, :m].T.astype(np.float64))
    Q = QtI.T
    back = np.linalg.norm(A - Q @ R) / (np.linalg.norm(A) * n * eps)
    orth = np.linalg.norm(QtI @ QtI.T - np.eye(m)) / (n * eps)
    return back, orth, R
