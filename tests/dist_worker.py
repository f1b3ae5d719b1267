"""Worker of the CPU multi-process test (gloo, world_size 2): the 1-D block-cyclic tiled QR with the
ORACLE's tile kernels standing in for the GPU kernels.  Each rank holds its own matrix copy, runs
the panels of its tile columns and the sweeps b200qr_sched_rank_tickets() gives it, and receives the
panel column (reflector tiles + tau) of every foreign step by broadcast -- exactly the data the
CUDA kernel reads over NVLink.  Rank 0 gathers the owned columns and compares with the
single-process oracle, bit for bit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class T(C.Structure):
    _fields_ = [("type", C.c_int), ("l", C.c_int), ("m", C.c_int), ("k", C.c_int), ("status", C.c_int)]


def run(rank, world, port, m, n, G, result_path):
    import torch
    import torch.distributed as dist
    import pyoracle as po
    import gpu_tiled_qr_decomposition_b200 as q
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, N = m // 32, n // 32
        K = min(M, N)
        L = po.lib()
        L.oracle_do_task.argtypes = [C.POINTER(T), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        work = np.zeros(64, dtype=np.float32)
        a = po.randzo(m, n)
        mat, tau = a.copy(), np.zeros_like(a)
        sweeps, panels = q.rank_tickets(M, N, G, rank, world)
        my_sweeps = {}
        for (k, g) in sweeps.tolist():
            my_sweeps.setdefault(k, []).append(g)
        # tickets of every rank: disjoint cover of the global streams, each in global order
        allsw = [None] * world
        dist.all_gather_object(allsw, (sweeps.tolist(), panels.tolist()))
        if rank == 0:
            glob = np.zeros((q.lib().b200qr_sched_sweep_order(M, N, G, None, 0), 2), dtype=np.int32)
            q.lib().b200qr_sched_sweep_order(M, N, G, glob.ctypes.data_as(C.c_void_p), len(glob))
            pos = {tuple(x): i for i, x in enumerate(glob.tolist())}
            seen = set()
            for r, (sw, pa) in enumerate(allsw):
                idx = [pos[tuple(x)] for x in sw]
                assert idx == sorted(idx), "rank stream is not a subsequence of the global order"
                assert all(g % world == r for _, g in sw) and all((k // G) % world == r for k in pa)
                assert not (seen & set(map(tuple, sw)))
                seen |= set(map(tuple, sw))
            assert seen == set(pos), "sweeps of all ranks do not cover the global stream"
            assert sorted(k for _, pa in allsw for k in pa) == list(range(K))

        def do(ty, l, mm, k):
            L.oracle_do_task(C.byref(T(ty, l, mm, k, 0)), mat.ctypes.data, tau.ctypes.data, 32, m, work.ctypes.data)

        for k in range(K):
            owner = (k // G) % world
            if rank == owner:
                assert k in panels
                do(0, k, k, k)
                for i in range(k + 1, M):
                    do(2, i, k, k)
            # the panel column of step k (reflector tiles + their tau) goes to everyone
            buf = torch.from_numpy(np.stack([mat[k * 32:(k + 1) * 32, :], tau[k * 32:(k + 1) * 32, :]]))
            dist.broadcast(buf, src=owner)
            if rank != owner:
                mat[k * 32:(k + 1) * 32, :] = buf[0].numpy()
                tau[k * 32:(k + 1) * 32, :] = buf[1].numpy()
            for g in my_sweeps.get(k, []):
                for j in range(max(g * G, k + 1), min((g + 1) * G, N)):
                    do(1, k, j, k)
                    for i in range(k + 1, M):
                        do(3, i, j, k)
        # gather the owned tile columns on rank 0
        outs = [torch.zeros(2, n, m) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(np.stack([mat, tau])), outs, dst=0)
        if rank == 0:
            out, t2 = np.zeros_like(a), np.zeros_like(a)
            for j in range(N):
                o = (j // G) % world
                out[j * 32:(j + 1) * 32] = outs[o][0].numpy()[j * 32:(j + 1) * 32]
                t2[j * 32:(j + 1) * 32] = outs[o][1].numpy()[j * 32:(j + 1) * 32]
            ref_out, ref_tau = po.tiled_qr(a, m, n)
            ok = bool(np.array_equal(out, ref_out) and np.array_equal(t2, ref_tau))
            with open(result_path, "w") as f:
                f.write("OK" if ok else "MISMATCH")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def run_tsqr(rank, world, port, m_local, n, result_path):
    """Multi-GPU TSQR exchange (TsqrDist) over gloo with the ORACLE's flat-tree tiled QR standing in for the CUDA
    engines: every rank reduces its own row block, the R factors are all-gathered, every rank reduces the stack.  All
    ranks must end with the same R, equal to float64 LAPACK's R of the whole matrix up to row signs."""
    import torch
    import torch.distributed as dist
    import pyoracle as po
    import gpu_tiled_qr_decomposition_b200 as q
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def oracle_reduce(rows):
            def fn(A, R):                                   # A: (n, rows) tensor = column-major rows x n
                out, _ = po.tiled_qr(np.ascontiguousarray(A.numpy()), rows, n)
                R.copy_(torch.from_numpy(np.ascontiguousarray(np.triu(out[:, :n].T).T)))   # [c, r], zero below the diagonal
            return fn
        rng = np.random.default_rng(1234)                   # same full matrix on every rank
        full = rng.uniform(-1, 1, size=(n, m_local * world)).astype(np.float32)
        A = torch.from_numpy(np.ascontiguousarray(full[:, rank * m_local:(rank + 1) * m_local]))
        R = torch.zeros((n, n), dtype=torch.float32)
        td = q.TsqrDist(m_local, n, world, local_reduce=oracle_reduce(m_local), top_reduce=oracle_reduce(world * n))
        td.factor(A, R)
        Rm = R.numpy().T.astype(np.float64)                 # (row, col)
        ok = np.count_nonzero(np.tril(Rm, -1)) == 0
        R64 = np.linalg.qr(full.T.astype(np.float64), mode="r")
        sgn = np.sign(np.diag(R64)) * np.sign(np.diag(Rm))
        ok = ok and np.max(np.abs(Rm - sgn[:, None] * R64)) <= 1e-4 * np.max(np.abs(R64))
        allR = [torch.empty_like(R) for _ in range(world)]
        dist.all_gather(allR, R)
        ok = ok and all(torch.equal(allR[0], x) for x in allR)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            with open(result_path, "w") as f:
                f.write("OK" if flag.item() == 1 else "MISMATCH")
    finally:
        dist.destroy_process_group()
