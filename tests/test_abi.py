"""C-ABI surface and host-side scheduler logic of libqrb200.so (no GPU needed).

 * the library loads and exports every function include/*.h declares;
 * the progress-word scheduler mirror (csrc/sched_host.cpp), driven by one serial worker with
   the reference's pick rule, emits gridscheduler.c's exact task sequence (golden traces),
   its READY set evolves like the reference scheduler's, predecessor sets and dependency
   counters equal SURVEY 8a row a7;
 * the persistent kernel's ticket order is a topological order of that DAG.
"""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        names += re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{}()]*\)\s*;", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(qr):
    L = qr.lib()
    names = declared_functions()
    assert "cudaQRTask" in names and "b200qr_factor_device" in names and len(names) >= 28
    for n in names:
        assert hasattr(L, n), "libqrb200.so does not export %s" % n
    assert set(names) == set(qr.EXPORTS), set(names) ^ set(qr.EXPORTS)
    assert L.b200qr_version() >= 100
    assert L.b200qr_strerror(0) == b"ok"


def test_invalid_arguments_are_rejected_without_a_gpu(qr):
    L = qr.lib()
    h = C.c_void_p()
    assert L.b200qr_plan_create(C.byref(h), 100, 64, None) == -1      # not a multiple of 32
    assert L.b200qr_plan_create(C.byref(h), 0, 64, None) == -1
    assert L.b200qr_factor_device(None, None, 64, None) == -1
    assert L.b200qr_tile_task(9, 0, 0, 0, None, None, 32) == -1
    o = qr.Opts()
    o.engine = 7
    assert L.b200qr_plan_create(C.byref(h), 64, 64, C.byref(o)) in (-1, -2)    # bad engine (-2: no device to ask first)
    # round-2 entry points: shapes and options are validated before anything touches the device
    assert L.b200qr_tiles_create(C.byref(h), 256, 256, 48) == -1              # tile size must be 32 / 64 / 128
    assert L.b200qr_tiles_create(C.byref(h), 256, 192, 128) == -1             # n % b
    assert L.b200qr_any_create(C.byref(h), 0, 5, None) == -1
    assert L.b200qr_tsqr_dist_create(C.byref(h), 4096, 256, 0, 9, 0) == -1    # at most 8 ranks
    assert L.b200qr_tsqr_dist_create(C.byref(h), 4096, 256, 2, 2, 0) == -1    # rank >= world
    assert L.b200qr_tsqr_dist_create(C.byref(h), 100, 256, 0, 1, 0) == -1     # fewer rows than columns
    assert L.b200qr_tsqr_create_tiles(C.byref(h), 4096, 288, 4) == -1         # n <= 256
    assert L.b200qr_tsqr_create_tiles(C.byref(h), 4096, 256, 3) == -1         # leaves must divide m into whole tile rows
    assert L.b200qr_form_q(None, 64, 64, 64, None, None, 64, 32, None) == -1
    assert L.b200qr_trsm_upper(None, 64, 64, None, 64, 1, None) == -1
    assert L.b200qr_solve_ls(None, 32, 64, 32, None, None, 32, 32, None) == -1  # m < n
    assert L.b200qr_dist_connect(None, None) == -1 and L.b200qr_tsqr_dist_connect(None, None) == -1


def test_no_cpu_fallback(qr):
    """On a machine without a GPU the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = np.zeros((64, 64), dtype=np.float32)
    with pytest.raises(qr.QRError):
        qr.factor_host(a, 64, 64)
    src = "".join(open(os.path.join(ROOT, "gpu_tiled_qr_decomposition_b200", f)).read()
                  for f in ("__init__.py", "csrc/qr_persistent.cu", "csrc/sched_host.cpp", "csrc/qr_batched.cu",
                            "csrc/qr_frontend.cu", "csrc/qr_tiles.cu", "csrc/qr_persistent_tc3_body.inc", "csrc/qr_device.cuh"))
    # the product never loads, links or imports the test oracle (comments may cite it as the parity reference)
    for needle in ("liboracle", "pyoracle", "oracle/", "_ref/", "libref", "dlopen"):
        assert needle not in src, needle


@pytest.mark.parametrize("key", ["1x1", "1x3", "3x1", "2x2", "3x3", "4x4", "7x3", "3x7", "16x4", "40x8", "32x32", "64x64"])
def test_scheduler_mirror_reproduces_reference_order(qr, golden, key):
    M, N = map(int, key.split("x"))
    tr = qr.sched_trace(M, N)
    rec = golden["sched_traces"][key]
    assert len(tr) == rec["tasks"] == sum((M - k) * (N - k) for k in range(min(M, N)))
    text = "".join("%d %d %d %d\n" % tuple(r) for r in tr)
    assert hashlib.sha256(text.encode()).hexdigest() == rec["sha256"]


def ref_ready_sets(po, M, N):
    """READY set after every completion, from the oracle's restatement of gridscheduler.c."""
    L = po.lib()

    class T(C.Structure):
        _fields_ = [("type", C.c_int), ("l", C.c_int), ("m", C.c_int), ("k", C.c_int), ("status", C.c_int)]

    class S(C.Structure):
        _fields_ = [("M", C.c_int), ("N", C.c_int), ("g", C.POINTER(T))]

    L.oracle_sched_init.restype = C.POINTER(S)
    L.oracle_sched_init.argtypes = [C.c_int, C.c_int]
    L.oracle_sched_next.argtypes = [C.POINTER(S), C.POINTER(T)]
    L.oracle_sched_done.argtypes = [C.POINTER(S), C.POINTER(T)]
    L.oracle_sched_free.argtypes = [C.POINTER(S)]
    s = L.oracle_sched_init(M, N)
    t = T()
    sets = []
    while L.oracle_sched_next(s, C.byref(t)) == 0:
        L.oracle_sched_done(s, C.byref(t))
        g = s.contents.g
        sets.append(sorted((g[y * M + x].type, x, y, g[y * M + x].k)
                           for x in range(M) for y in range(N) if g[y * M + x].status == 0))
    L.oracle_sched_free(s)
    return sets


@pytest.mark.parametrize("M,N", [(1, 1), (3, 3), (5, 5), (7, 3), (3, 7), (9, 6)])
def test_ready_set_evolution_equals_reference(qr, po, M, N):
    want = ref_ready_sets(po, M, N)
    L = qr.lib()
    s = L.b200qr_sched_init(M, N)
    t = qr.Task()
    got = []
    ready = {(0, 0, 0, 0)}
    while L.b200qr_sched_next(s, C.byref(t)) == qr.TASK_AVAIL:
        ready.discard((t.type, t.l, t.m, t.k))
        L.b200qr_sched_done(s, C.byref(t))
        # reconstruct the READY set from the progress words: a tile is READY iff its next task's
        # inputs are final and it is not running (single worker => nothing is running here)
        cur = set()
        for l in range(M):
            for m in range(N):
                k = L.b200qr_sched_progress(s, l, m)
                if k > min(l, m):
                    continue
                if l > k and L.b200qr_sched_progress(s, l - 1, m) < k + 1:
                    continue
                if m > k and L.b200qr_sched_progress(s, l, k) < k + 1:
                    continue
                ty = (qr.QRS if m == k else qr.SAPP) if l == k else (qr.QRD if m == k else qr.DAPP)
                cur.add((ty, l, m, k))
        assert len(cur) == L.b200qr_sched_ready_count(s)
        got.append(sorted(cur))
    L.b200qr_sched_free(s)
    assert got == want


def predecessors_by_definition(ty, l, m, k):
    """SURVEY 8a row a7 (gridscheduler.c:112-174)."""
    def typ(x, y):
        return (0 if y == k else 1) if x == k else (2 if y == k else 3)
    pred = set()
    if k > 0:
        pred.add((3, l, m, k - 1))
    if ty == 1:
        pred.add((0, k, k, k))
    if ty == 2:
        pred.add((typ(l - 1, k), l - 1, k, k))
    if ty == 3:
        pred.add((2, l, k, k))
        pred.add((typ(l - 1, m), l - 1, m, k))
    return pred


def test_predecessor_sets_and_dependency_counters(qr):
    L = qr.lib()
    M, N = 6, 5
    for k in range(min(M, N)):
        for l in range(k, M):
            for m in range(k, N):
                ty = (0 if m == k else 1) if l == k else (2 if m == k else 3)
                t = qr.Task(ty, l, m, k)
                buf = (qr.Task * 3)()
                n = L.b200qr_sched_predecessors(C.byref(t), buf)
                got = {(buf[i].type, buf[i].l, buf[i].m, buf[i].k) for i in range(n)}
                assert got == predecessors_by_definition(ty, l, m, k)
                assert n == L.b200qr_sched_dep_count(ty, k)
    # counters {QRS: k>0, SAPP: 1+(k>0), QRD: 1+(k>0), DAPP: 2+(k>0)}
    assert [L.b200qr_sched_dep_count(t, 0) for t in range(4)] == [0, 1, 1, 2]
    assert [L.b200qr_sched_dep_count(t, 3) for t in range(4)] == [1, 2, 2, 3]


@pytest.mark.parametrize("M,N", [(1, 1), (2, 2), (5, 5), (9, 4), (4, 9), (13, 13), (40, 8), (33, 37)])
def test_device_ticket_order_is_a_topological_order(qr, M, N):
    L = qr.lib()
    order = qr.device_order(M, N)
    pos = {tuple(r): i for i, r in enumerate(order.tolist())}
    assert len(pos) == len(order) == sum((M - k) * (N - k) for k in range(min(M, N)))
    for (ty, l, m, k), i in pos.items():
        t = qr.Task(ty, l, m, k)
        buf = (qr.Task * 3)()
        n = L.b200qr_sched_predecessors(C.byref(t), buf)
        for j in range(n):
            p = (buf[j].type, buf[j].l, buf[j].m, buf[j].k)
            assert pos[p] < i, (p, (ty, l, m, k))


def test_lagged_sweep_order_is_topological(qr):
    """b200qr_sched_sweep_order_lag: same sweeps as the skewed order; for a fixed group steps ascend; every S(k, .) comes
    after S(k-1, diagonal group of step k-1) -- the sweep P(k) waits for."""
    import ctypes as C
    L = qr.lib()
    for (M, N, G, LAG) in ((24, 24, 4, 6), (40, 28, 4, 16), (16, 33, 4, 0), (64, 64, 4, 16)):
        n = L.b200qr_sched_sweep_order_lag(M, N, G, LAG, None, 0)
        assert n == L.b200qr_sched_sweep_order(M, N, G, None, 0)
        out = np.zeros((n, 2), dtype=np.int32)
        assert L.b200qr_sched_sweep_order_lag(M, N, G, LAG, out.ctypes.data_as(C.c_void_p), n) == n
        ref = np.zeros((n, 2), dtype=np.int32)
        L.b200qr_sched_sweep_order(M, N, G, ref.ctypes.data_as(C.c_void_p), n)
        assert sorted(map(tuple, out)) == sorted(map(tuple, ref))
        pos = {tuple(kg): i for i, kg in enumerate(map(tuple, out))}
        for (k, g), i in pos.items():
            if (k - 1, g) in pos:
                assert pos[(k - 1, g)] < i
            if k >= 1 and (k - 1, k // G) in pos:                  # diagonal group of step k-1 holds tile column k
                assert pos[(k - 1, k // G)] < i


def test_final_progress_words_match_the_host_mirror(qr):
    """The words the overlapped cudaQRTask path waits for (b200qr_sched_final_progress) are the host mirror's progress
    words once the DAG has run dry, plus the panel's "column final" publish on the diagonal."""
    import ctypes as C
    L = qr.lib()
    T = qr.Task
    for (p, q) in [(1, 1), (4, 4), (7, 3), (3, 7), (16, 4), (9, 9)]:
        s = L.b200qr_sched_init(p, q)
        t = T()
        while L.b200qr_sched_next(s, C.byref(t)) == qr.TASK_AVAIL:
            L.b200qr_sched_done(s, C.byref(t))
        K = min(p, q)
        for l in range(p):
            for m in range(q):
                want = L.b200qr_sched_progress(s, l, m) + (1 if (l == m and m < K) else 0)
                assert L.b200qr_sched_final_progress(p, q, l, m) == want, (p, q, l, m)
        L.b200qr_sched_free(s)
        assert L.b200qr_sched_final_progress(p, q, p, 0) == -1 and L.b200qr_sched_final_progress(p, q, 0, -1) == -1
