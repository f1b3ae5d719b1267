"""Multi-rank path on ONE GPU: two rank objects share the device and run in a single cooperative
launch (b200qr_dist_factor_group), each reading the other's reflector tiles / tau / progress words
through the same peer pointers a real 2-GPU run uses.  The result must be bit-identical to the
single-GPU factorisation (every tile sees the same operation sequence) and match the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gather(ranks, m, n):
    """Assemble the factored matrix and tau (reference layout) from the owners."""
    p, q = m // 32, n // 32
    K = min(p, q)
    out = np.zeros((n, m), dtype=np.float32)
    tau = np.zeros((n, m), dtype=np.float32)
    for rk in ranks:
        a = np.zeros((n, m), dtype=np.float32)
        tc = np.zeros((K, p, 32), dtype=np.float32)
        rk.download(a.ctypes.data)
        rk.download_tau(tc.ctypes.data)
        for j in rk.owned_tile_cols():
            out[j * 32:(j + 1) * 32] = a[j * 32:(j + 1) * 32]
            if j < K:
                for i in range(j, p):
                    tau[j * 32, i * 32:(i + 1) * 32] = tc[j, i]
    return out, tau


@pytest.mark.parametrize("m,n", [(512, 512), (1024, 768), (2048, 2048)])
def test_two_ranks_on_one_device_match_single_gpu(cuda, qr, po, m, n):
    torch = cuda
    a = po.randzo(m, n)
    ranks = [qr.DistRank(m, n, r, 2, share_device=2) for r in range(2)]
    qr.dist_connect_local(ranks)
    for rk in ranks:
        rk.upload(a.ctypes.data)
    qr.dist_factor_group(ranks, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out, tau = gather(ranks, m, n)
    cols = [set(rk.owned_tile_cols()) for rk in ranks]
    assert ranks[0].block in (8, 16) and cols[0] | cols[1] == set(range(n // 32)) and not (cols[0] & cols[1])
    # single-GPU result
    d = torch.from_numpy(a).cuda()
    dt = torch.zeros_like(d)
    plan = qr.Plan(m, n)
    s = torch.cuda.current_stream().cuda_stream
    plan.factor(d.data_ptr(), m, s)
    plan.tau_to_reference(dt.data_ptr(), m, s)
    torch.cuda.synchronize()
    assert np.array_equal(out, d.cpu().numpy())
    assert np.array_equal(tau, dt.cpu().numpy())
    plan.close()
    ref_out, ref_tau = po.tiled_qr(a, m, n)
    scale = float(np.max(np.abs(ref_out)))
    assert np.max(np.abs(out - ref_out)) <= 1e-4 * max(1.0, scale)
    for rk in ranks:
        rk.close()


@pytest.mark.parametrize("m,n", [(2048, 2048), (4096, 3072)])
def test_tensor_core_engine_two_ranks_on_one_device(cuda, qr, po, m, n):
    """The multi-GPU form of the tensor-core engine (b200qr_opts.engine = 2: operand tiles TMA-stored into EVERY rank's
    workspace by the panel owner, progress words pushed into every rank's array) with both ranks on one device in a
    single cooperative launch -- the same code path a 2-GPU run takes, peer pointers included, runnable on a 1-GPU
    box.  The result must equal the single-GPU engine's bit for bit and the oracle within the engine's tolerance."""
    torch = cuda
    a = po.randzo(m, n)
    ranks = [qr.DistRank(m, n, r, 2, share_device=2, engine=2) for r in range(2)]
    assert all(rk.variant == 5 for rk in ranks) and ranks[0].block == 4
    qr.dist_connect_local(ranks)
    for rk in ranks:
        rk.upload(a.ctypes.data)
    for rep in range(2):                               # twice: the second run starts from the first one's final words
        for rk in ranks:
            rk.upload(a.ctypes.data)
        qr.dist_factor_group(ranks, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    out, tau = gather(ranks, m, n)
    d = torch.from_numpy(a).cuda()
    dt = torch.zeros_like(d)
    plan = qr.Plan(m, n, engine=2)
    s = torch.cuda.current_stream().cuda_stream
    plan.factor(d.data_ptr(), m, s)
    plan.tau_to_reference(dt.data_ptr(), m, s)
    torch.cuda.synchronize()
    assert np.array_equal(out, d.cpu().numpy())
    assert np.array_equal(tau, dt.cpu().numpy())
    plan.close()
    ref_out, _ = po.tiled_qr(a, m, n)
    from qr_checks import r_rel_err_up_to_row_sign
    err, flips = r_rel_err_up_to_row_sign(out, ref_out, m, n)
    assert err <= 5e-4 and flips <= 16, (err, flips)
    for rk in ranks:
        rk.close()


def test_ranks_must_agree_on_the_plan(cuda, qr):
    """b200qr_dist_connect refuses peers whose plan differs (kernel variant, distribution block, shape): the exported
    handle blob carries those facts (a rank that chose another engine would own other columns)."""
    import ctypes as C
    a = qr.DistRank(2048, 2048, 0, 2, share_device=2, engine=2)
    b = qr.DistRank(2048, 2048, 1, 2, share_device=2, engine=1)
    assert a.variant == 5 and b.variant == 2
    blobs = [a.export(), b.export()]
    for rk in (a, b):
        rc = qr.lib().b200qr_dist_connect(rk._h, C.c_char_p(b"".join(blobs)))
        assert rc == -1, rc                              # B200QR_EINVAL, in both directions
    a.close(); b.close()


@pytest.mark.parametrize("m,n,packed,engine", [(2048, 2048, False, 0), (4096, 3072, True, 0), (4096, 4096, True, 2), (1024, 4096, False, 0)])
def test_rank_factor_host_overlapped_matches_device_path(cuda, qr, m, n, packed, engine):
    """b200qr_dist_reset_host / b200qr_dist_factor_host (host data, copies overlapped with the kernel) on a one-rank
    world: the matrix that comes back must be the device path's, bit for bit, in both host layouts."""
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1      # row of the tensor = matrix column
    plan = qr.Plan(m, n, engine=engine)
    ref = A.clone()
    s = torch.cuda.current_stream().cuda_stream
    plan.factor(ref.data_ptr(), m, s)
    torch.cuda.synchronize()
    rk = qr.DistRank(m, n, 0, 1, engine=engine)
    assert rk.variant == plan.variant
    ldh = m + 32                                                       # a leading dimension of its own
    for rep in range(2):                                               # the second call reuses streams and scratch
        h = torch.zeros((n, ldh), dtype=torch.float32).pin_memory()
        h[:, :m] = A.cpu()
        rk.reset_host()
        ms = rk.factor_host(h.data_ptr(), ldh, packed=packed)          # one rank owns everything: both layouts coincide
        assert ms > 0
        assert torch.equal(h[:, :m], ref.cpu())
        assert float(h[:, m:].abs().max()) == 0.0                      # the padding of the host matrix is not touched
    # protocol: factor_host without reset_host is refused
    assert qr.lib().b200qr_dist_factor_host(rk._h, h.data_ptr(), ldh, 0, None) == -1
    rk.close(); plan.close()


def test_tensor_core_engine_on_two_real_gpus(cuda):
    """The multi-GPU form of the tensor-core engine (operand tiles broadcast by TMA stores into every rank's workspace)
    must give the single-GPU engine's result bit for bit.  Needs two visible GPUs (skipped on a one-GPU box)."""
    import os
    import subprocess
    import sys
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QRB200_WATCHDOG_MS="30000")
    env.pop("QRB200_TC", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tools", "dist_tc3_check.py"), "12288"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "single-GPU variant 5" in r.stdout and "identical: True" in r.stdout, r.stdout[-2000:]
    assert "host identical: True" in r.stdout, r.stdout[-2000:]      # b200qr_dist_factor_host, copies overlapped


def test_simt_engine_on_two_real_gpus(cuda):
    """The SIMT engine's multi-GPU form (sweeps read the panel owner's reflector tiles, tau and progress words over
    NVLink inside the kernel) on two real GPUs: bit-identical to one GPU, also through the overlapped host path."""
    import os
    import subprocess
    import sys
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QRB200_WATCHDOG_MS="30000", QRB200_TC="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29535",
                        os.path.join(root, "tools", "dist_tc3_check.py"), "4096"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "single-GPU variant 2" in r.stdout and "identical: True" in r.stdout, r.stdout[-2000:]
    assert "host identical: True" in r.stdout, r.stdout[-2000:]


def test_tsqr_split_over_two_real_gpus(cuda):
    """configs[3] on two GPUs (TsqrDist: local streaming reduction, NCCL all-gather of the R factors, redundant top
    reduction): same R on both ranks, R^T R = A^T A.  Needs two visible GPUs (skipped on a one-GPU box)."""
    import json
    import os
    import subprocess
    import sys
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "tools", "bench_tsqr_dist.py"), "262144", "256"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["same_R_on_all_ranks"] and line["gram_rel_err"] <= 50 * float(np.finfo(np.float32).eps), line


@pytest.mark.parametrize("world,mode", [(2, "stream"), (4, "tiles"), (3, "stream")])
def test_tsqr_dist_c_abi_ranks_on_one_device(cuda, qr, world, mode):
    """b200qr_tsqr_dist_* (the multi-GPU TSQR behind the C ABI: R factors exchanged over peer memory, epoch flags with
    system-scope release / acquire, two alternating stack buffers) with all rank objects on one device and one stream:
    every rank ends with the same R, R^T R = sum_r A_r^T A_r, also on the second and third call (epochs, buffer parity)."""
    torch = cuda
    m_local, n = 16384, 128
    s = torch.cuda.current_stream().cuda_stream
    ranks = [qr.TsqrDistC(m_local, n, r, world, mode=mode) for r in range(world)]
    qr.tsqr_dist_connect_local(ranks)
    g = torch.Generator(device="cuda").manual_seed(41)
    for rep in range(3):
        A = [torch.rand((n, m_local), generator=g, device="cuda") * 2 - 1 for _ in range(world)]
        d = [a.clone() for a in A]
        R = [torch.full((n, n), 3.0, device="cuda") for _ in range(world)]
        for r in range(world):
            ranks[r].factor_begin(d[r].data_ptr(), m_local, s)
        for r in range(world):
            ranks[r].factor_end(R[r].data_ptr(), s)
        torch.cuda.synchronize()
        assert all(torch.equal(R[0], x) for x in R)
        Rm = R[0].t().double()
        assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
        G = sum(a.double() @ a.double().t() for a in A)
        assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    for rk in ranks:
        rk.close()


def test_tsqr_dist_single_rank(cuda, qr):
    """TsqrDist with one rank = the plain streaming plan."""
    torch = cuda
    m, n = 65536, 128
    g = torch.Generator(device="cuda").manual_seed(9)
    A0 = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    A, R = A0.clone(), torch.zeros((n, n), device="cuda")
    td = qr.TsqrDist(m, n, 1)
    td.factor(A, R)
    torch.cuda.synchronize()
    td.close()
    Rm, Ad = R.t().double(), A0.t().double()
    G = Ad.t() @ Ad
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
