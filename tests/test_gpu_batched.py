"""Batched small QR (BASELINE configs[2]) and TSQR (configs[3]) against the oracle / float64 LAPACK."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def batched(qr, torch, mats, m, n):
    """mats: (nprob, n, m) float32 (each problem column-major m x n, ld = m); returns (out, tau_compact)."""
    nprob = mats.shape[0]
    d = torch.from_numpy(np.ascontiguousarray(mats)).cuda()
    fl = C.c_longlong(0)
    qr.check(qr.lib().b200qr_batched_workspace_floats(m, n, nprob, C.byref(fl)))
    tau = torch.zeros(fl.value, dtype=torch.float32, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    qr.check(qr.lib().b200qr_batched_factor(C.c_void_p(d.data_ptr()), m, n, m, m * n, nprob, C.c_void_p(tau.data_ptr()),
                                            C.c_void_p(ctr.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return d.cpu().numpy(), tau.cpu().numpy()


def expand_tau(tc, m, n):
    p, q = m // 32, n // 32
    K = min(p, q)
    t = tc.reshape(K, p, 32)
    out = np.zeros((n, m), dtype=np.float32)
    for k in range(K):
        for i in range(k, p):
            out[k * 32, i * 32:(i + 1) * 32] = t[k, i]
    return out


@pytest.mark.parametrize("m,n,nprob", [(256, 256, 5), (32, 32, 3), (64, 256, 2), (512, 128, 4), (2048, 256, 2)])
def test_batched_matches_oracle_per_problem(cuda, qr, po, m, n, nprob):
    mats = np.stack([po.randzo(m, n, seed=5 + b) for b in range(nprob)])
    out, tau = batched(qr, cuda, mats, m, n)
    per = tau.size // nprob
    for b in range(nprob):
        ref_out, ref_tau = po.tiled_qr(mats[b], m, n)
        scale = max(1.0, float(np.max(np.abs(ref_out))))
        assert np.max(np.abs(out[b] - ref_out)) <= 1e-4 * scale
        assert np.max(np.abs(expand_tau(tau[b * per:(b + 1) * per], m, n) - ref_tau)) <= 1e-5 * max(1, n // 32)


def test_batched_equals_single_matrix_path(cuda, qr, po):
    """A 256x256 problem of the batch is bit-identical to b200qr_factor_device on that matrix."""
    torch = cuda
    m = n = 256
    mats = np.stack([po.randzo(m, n, seed=9 + b) for b in range(300)])       # more problems than CTAs
    out, _ = batched(qr, torch, mats, m, n)
    plan = qr.Plan(m, n)
    for b in (0, 7, 299):
        d = torch.from_numpy(mats[b]).cuda()
        plan.factor(d.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(out[b], d.cpu().numpy())
    plan.close()


@pytest.mark.parametrize("m,n,nprob", [(256, 256, 1000), (512, 96, 700), (64, 160, 1203), (32, 32, 400), (96, 256, 301)])
def test_batched_one_warp_per_problem_kernel(cuda, qr, po, monkeypatch, m, n, nprob):
    """Enough problems (>= 4 per SM) select the round-2 kernel -- one warp per problem, panel and update phases in
    lock-step across the CTA, the next tile of every chain prefetched (batched_warp_kernel).  Every problem must be bit-identical to
    b200qr_factor_device on that matrix (same tile kernels, same operation order), which the oracle tests pin; a sample
    is also compared with the oracle directly.  Odd tile-column counts exercise the single-column tail of the update."""
    torch = cuda
    monkeypatch.setenv("QRB200_BATCHED", "warp")
    rng = np.random.default_rng(m + n + nprob)
    mats = rng.uniform(-1, 1, (nprob, n, m)).astype(np.float32)
    mats[3, 5:9, :] = 0.0                                              # a problem with zero columns (degenerate reflectors)
    out, tau = batched(qr, torch, mats, m, n)
    per = tau.size // nprob
    plan = qr.Plan(m, n)
    s = torch.cuda.current_stream().cuda_stream
    dt = torch.zeros((n, m), device="cuda")
    for b in (0, 3, 7, 8, 9, nprob // 2, nprob - 2, nprob - 1):
        d = torch.from_numpy(mats[b]).cuda()
        plan.factor(d.data_ptr(), m, s)
        plan.tau_to_reference(dt.data_ptr(), m, s)
        torch.cuda.synchronize()
        assert np.array_equal(out[b], d.cpu().numpy()), b
        assert np.array_equal(expand_tau(tau[b * per:(b + 1) * per], m, n), dt.cpu().numpy()), b
    plan.close()
    for b in (1, nprob - 1):
        ref_out, ref_tau = po.tiled_qr(mats[b], m, n)
        assert np.max(np.abs(out[b] - ref_out)) <= 1e-4 * max(1.0, float(np.max(np.abs(ref_out))))
    # every problem: R^T R = A^T A
    R = np.triu(np.transpose(out, (0, 2, 1))[:, :min(m, n), :].astype(np.float64))
    A = np.transpose(mats, (0, 2, 1)).astype(np.float64)
    G = np.einsum("bij,bik->bjk", A, A)
    RtR = np.einsum("bij,bik->bjk", R, R)
    err = np.linalg.norm((RtR - G).reshape(nprob, -1), axis=1) / np.maximum(np.linalg.norm(G.reshape(nprob, -1), axis=1), 1e-30)
    assert float(err.max()) <= 50 * float(np.finfo(np.float32).eps), float(err.max())


@pytest.mark.parametrize("m,n,leaves", [(1 << 20, 256, 0), (65536, 128, 37 * 0 + 64), (40960, 256, 5), (8192, 32, 256)])
def test_tsqr_tile_engine_r_factor(cuda, qr, m, n, leaves):
    """TSQR with the leaves on the batched tile kernel (b200qr_tsqr_create_tiles: one warp per leaf, the reference's
    flat TSQRT chain per leaf) and pairwise triangular levels above: R upper triangular, R^T R = A^T A, R equal to
    float64 LAPACK's up to row signs within 1e-4 max|R|."""
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(17)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = A.clone()
    t = qr.Tsqr(m, n, strips=leaves, mode="tiles")
    R = torch.full((n, n), 7.0, dtype=torch.float32, device="cuda")
    for _ in range(2):
        d.copy_(A)
        t.factor(d.data_ptr(), m, R.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert "tile engine" in t.describe()
    t.close()
    Rm = R.t().double()
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    Ad = A.t().double()
    G = Ad.t() @ Ad
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    Rc = torch.linalg.cholesky(G).t()
    sg = torch.sign(torch.diagonal(Rm))
    assert (torch.max(torch.abs(Rm - sg[:, None] * Rc)) / torch.max(torch.abs(Rc))).item() <= 1e-4
    assert torch.isfinite(d).all()


@pytest.mark.parametrize("m,n,leaf,fanin", [(4096, 256, 512, 2), (8192, 128, 256, 4), (12288, 256, 2048, 4), (2048, 64, 2048, 2),
                                              (5120, 256, 512, 4)])
def test_tsqr_r_factor(cuda, qr, m, n, leaf, fanin):
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1    # column-major m x n
    d = A.clone()
    h = C.c_void_p()
    qr.check(qr.lib().b200qr_tsqr_create(C.byref(h), m, n, leaf, fanin))
    R = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    qr.check(qr.lib().b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    qr.lib().b200qr_tsqr_destroy(h)
    Rm = R.t().double()                                  # (n, n) values, upper triangular
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    Ad = A.t().double()
    G = Ad.t() @ Ad
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    R64 = torch.linalg.qr(Ad, mode="r").R
    s = torch.sign(torch.diagonal(R64)) * torch.sign(torch.diagonal(Rm))
    assert (torch.max(torch.abs(Rm - s[:, None] * R64)) / torch.max(torch.abs(R64))).item() <= 1e-4


@pytest.mark.parametrize("m,n,strips", [(4096, 256, 0), (8192, 128, 3), (12288, 256, 5), (2048, 64, 1), (5120, 256, 0),
                                          (262144, 256, 0), (4128, 32, 2)])
def test_tsqr_streaming_r_factor(cuda, qr, m, n, strips):
    """Streaming TSQR engine (strips of rows, R resident in shared memory, tree over the strips' R factors):
    same contract as the leaf-batch tree -- R upper triangular, R^T R = A^T A, R = LAPACK's up to row signs."""
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1    # column-major m x n
    d = A.clone()
    h = C.c_void_p()
    qr.check(qr.lib().b200qr_tsqr_create_streaming(C.byref(h), m, n, strips))
    R = torch.full((n, n), 7.0, dtype=torch.float32, device="cuda")
    for _ in range(2):                                   # a plan is reusable
        d.copy_(A)
        qr.check(qr.lib().b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert qr.lib().b200qr_tsqr_levels(h) >= 1
    qr.lib().b200qr_tsqr_destroy(h)
    Rm = R.t().double()
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    Ad = A.t().double()
    G = Ad.t() @ Ad
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    R64 = torch.linalg.qr(Ad, mode="r").R
    s = torch.sign(torch.diagonal(R64)) * torch.sign(torch.diagonal(Rm))
    assert (torch.max(torch.abs(Rm - s[:, None] * R64)) / torch.max(torch.abs(R64))).item() <= 1e-4
    # the reflectors left in place are bounded (|v| <= 1 elementwise for a Householder vector scaled by its pivot)
    assert torch.isfinite(d).all() and d.abs().max().item() <= 1.0 + 1e-6


def test_tsqr_streaming_reflectors_reproduce_r(cuda, qr):
    """One strip: the reflector blocks left in place ARE the implicit Q.  Replaying them in float64 on [0; A] (the
    TSQRT recurrence starts from an empty triangle on top of the strip: v_c = [e_c; V2(:, c)] per block of rows,
    tau_c = 2 / (1 + |V2(:, c)|^2), qrdecomp.c:660-664) must give [R; 0] with the R the engine returned."""
    torch = cuda
    m, n = 1536, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = A.clone()
    h = C.c_void_p()
    qr.check(qr.lib().b200qr_tsqr_create_streaming(C.byref(h), m, n, 1))
    rb = qr.lib().b200qr_tsqr_stream_block_rows(h)
    assert rb > 0 and m % rb == 0
    R = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    qr.check(qr.lib().b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    qr.lib().b200qr_tsqr_destroy(h)
    V = d.t().double()                                   # (m, n): block rows of reflectors
    top = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    bot = A.t().double().clone()                         # (m, n)
    for blk in range(m // rb):
        rows = slice(blk * rb, (blk + 1) * rb)
        for c in range(n):
            v2 = V[rows, c]
            tau = 2.0 / (1.0 + torch.dot(v2, v2))
            w = tau * (top[c, :] + v2 @ bot[rows, :])
            top[c, :] -= w
            bot[rows, :] -= torch.outer(v2, w)
    scale = top.abs().max().item()
    assert bot.abs().max().item() <= 2e-5 * scale        # everything below the triangle is annihilated
    assert (top - R.t().double()).abs().max().item() <= 2e-5 * scale


@pytest.mark.parametrize("count,n", [(1, 64), (2, 256), (5, 128), (8, 256), (13, 32)])
def test_tsqr_rstack_reduction(cuda, qr, count, n):
    """Pairwise reduction of stacked upper-triangular factors (the top of the multi-GPU TSQR): R^T R = sum_i R_i^T R_i and
    R equals LAPACK's R of the stack up to row signs; the identity steps of the triangular blocks leave tau = 0."""
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(31 + count)
    blocks = torch.triu(torch.rand((count, n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1)   # [i, row, col]
    stack = blocks.reshape(count * n, n).t().contiguous()            # column-major (count*n) x n: [col, row]
    ref = stack.clone()
    t = qr.Tsqr(count * n, n, rstack=count)
    R = torch.full((n, n), 3.0, device="cuda")
    t.factor(stack.data_ptr(), count * n, R.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t.close()
    Rm = R.t().double()
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    S = ref.t().double()
    G = S.t() @ S
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    R64 = torch.linalg.qr(S, mode="r").R
    sg = torch.sign(torch.diagonal(R64)) * torch.sign(torch.diagonal(Rm))
    assert (torch.max(torch.abs(Rm - sg[:, None] * R64)) / torch.max(torch.abs(R64))).item() <= 1e-4


@pytest.mark.parametrize("m,n", [(1536, 64), (768, 256), (1536, 256), (1440, 224)])
def test_tsqr_streaming_matches_oracle_elementwise(cuda, qr, po, m, n):
    """One strip of the streaming engine against the oracle's chain of n-wide TSQRT steps (pyoracle.tsqr_stream = the
    reference's STSQRF restated, qrdecomp.c:689-721) on the same input and block height: R, its diagonal signs and EVERY
    reflector block agree element by element (float32 rounding only: FMA, split sums, tau = 1 + |x0| / nrm vs
    2 / (1 + |v|^2)), at BASELINE.json's width n = 256 too.  Round 1 could not pin n = 256: a strip starts from an empty
    triangle, and while fewer than n rows have gone by the trailing columns of a block are rounding noise whose signs
    decided the final row signs.  Both implementations now stop a start-up block at the last column its rows determine
    (oracle_tsqrt_lead; tsqr_strip_kernel's c_end), which is what those columns are mathematically: zero."""
    torch = cuda
    rng = np.random.default_rng(5 * m + n)
    a = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    h = C.c_void_p()
    qr.check(qr.lib().b200qr_tsqr_create_streaming(C.byref(h), m, n, 1))
    rb = qr.lib().b200qr_tsqr_stream_block_rows(h)
    assert m % rb == 0
    d = torch.from_numpy(a).cuda()
    R = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    qr.check(qr.lib().b200qr_tsqr_factor(h, C.c_void_p(d.data_ptr()), m, C.c_void_p(R.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    qr.lib().b200qr_tsqr_destroy(h)
    V_ref, _, R_ref = po.tsqr_stream(a, m, n, rb)
    scale = float(np.max(np.abs(R_ref)))
    assert np.max(np.abs(R.cpu().numpy() - R_ref)) <= 1e-4 * scale
    assert np.all(np.sign(np.diag(R.cpu().numpy())) == np.sign(np.diag(R_ref)))      # same sign convention
    assert np.max(np.abs(d.cpu().numpy() - V_ref)) <= 1e-4       # every reflector block, start-up blocks included
    if n > rb:                                                   # the undetermined columns of the first block are exactly zero
        assert not d.cpu().numpy()[rb:, :rb].any()


def test_tsqr_full_size(cuda, qr):
    """BASELINE.json configs[3] at full size, 4 194 304 x 256 on one GPU: R^T R = A^T A to 50 eps, R upper triangular,
    and R equal up to row signs to the Cholesky factor of the float64 Gram matrix (= the R of a QR of A, unique up to row
    signs; cond(A) ~ 1 for this input, so the float64 Cholesky is exact to ~1e-13) within 1e-4 max|R|."""
    torch = cuda
    m, n = 4194304, 256
    free, _ = torch.cuda.mem_get_info()
    if free < 3 * 4 * m * n:
        pytest.skip("needs 12 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = A.clone()
    t = qr.Tsqr(m, n)
    R = torch.zeros((n, n), device="cuda")
    t.factor(d.data_ptr(), m, R.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t.close()
    del d
    G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for c in range(0, m, 1 << 20):
        blk = A[:, c:c + (1 << 20)].double()
        G += blk @ blk.t()
    Rm = R.t().double()
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    assert (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item() <= 50 * float(np.finfo(np.float32).eps)
    Rc = torch.linalg.cholesky(G).t()                            # upper, positive diagonal
    s = torch.sign(torch.diagonal(Rm))
    assert (s != 0).all()
    assert (torch.max(torch.abs(Rm - s[:, None] * Rc)) / torch.max(torch.abs(Rc))).item() <= 1e-4
