"""form-Q, least squares and the arbitrary-size front-end (csrc/qr_frontend.cu), against
float64 LAPACK (torch.linalg) on the same inputs.  The reference has none of these (qrdecomp.c:1292-1311 multAB is an
unused stub; qrdecomp.c:162 / gpucalc.cu:1582 truncate m, n to multiples of 32).  Stated tolerances: orthogonality and
backward error <= 5 n eps (the SIMT engine's bound of the parity tests), solutions within 1e-4 relative of LAPACK's for
these well-conditioned inputs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float32).eps)


def _stream(torch):
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("m,n", [(512, 256), (1024, 1024), (2048, 512)])
def test_form_q(cuda, qr, m, n):
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(m + n)
    A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    d = A.clone()
    plan = qr.Plan(m, n)
    plan.factor(d.data_ptr(), m, _stream(torch))
    Q = torch.full((n, m), 9.0, device="cuda")                       # thin Q: m x n, column-major
    qr.check(qr.lib().b200qr_form_q(C.c_void_p(d.data_ptr()), m, n, m, C.c_void_p(plan.tau_ptr), C.c_void_p(Q.data_ptr()), m, n,
                                    C.c_void_p(_stream(torch))), "b200qr_form_q")
    torch.cuda.synchronize()
    Qm = Q.t().double()                                              # (m, n)
    orth = (torch.linalg.norm(Qm.t() @ Qm - torch.eye(n, device="cuda", dtype=torch.float64)) / (n * EPS)).item()
    assert orth <= 5.0, orth
    R = torch.triu(d.t()[:n].double())
    back = (torch.linalg.norm(Qm @ R - A.t().double()) / (torch.linalg.norm(A.double()) * n * EPS)).item()
    assert back <= 5.0, back
    plan.close()


@pytest.mark.parametrize("m,n,nrhs", [(1024, 512, 32), (2048, 2048, 64), (4096, 256, 96)])
def test_solve_ls(cuda, qr, m, n, nrhs):
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(7 * m + n)
    A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    B = torch.rand((nrhs, m), generator=g, device="cuda") * 2 - 1
    d, b = A.clone(), B.clone()
    plan = qr.Plan(m, n)
    plan.factor(d.data_ptr(), m, _stream(torch))
    qr.check(qr.lib().b200qr_solve_ls(C.c_void_p(d.data_ptr()), m, n, m, C.c_void_p(plan.tau_ptr), C.c_void_p(b.data_ptr()), m, nrhs,
                                      C.c_void_p(_stream(torch))), "b200qr_solve_ls")
    torch.cuda.synchronize()
    X = b[:, :n].t().double()                                        # (n, nrhs)
    want = torch.linalg.lstsq(A.t().double(), B.t().double()).solution
    rel = (torch.linalg.norm(X - want) / torch.linalg.norm(want)).item()
    cond = torch.linalg.cond(A.t().double()).item()
    assert rel <= max(1e-4, 20 * cond * EPS), (rel, cond)
    # normal equations: A^T (A x - b) = 0
    res = A.double() @ (A.t().double() @ X - B.t().double())
    assert (torch.linalg.norm(res) / (torch.linalg.norm(A.double()) ** 2 * torch.linalg.norm(X))).item() <= 50 * EPS
    plan.close()


@pytest.mark.parametrize("m,n", [(1000, 300), (257, 255), (33, 31), (700, 700), (2050, 130), (100, 37), (150, 260)])
def test_arbitrary_sizes(cuda, qr, m, n):
    """m, n not multiples of 32: R up to row signs, Q^T Q = I, A = Q R, least squares -- all for the REAL sizes."""
    torch = cuda
    g = torch.Generator(device="cuda").manual_seed(3 * m + n)
    ld = m + 3                                                      # unaligned leading dimension on purpose
    buf = torch.rand((n, ld), generator=g, device="cuda") * 2 - 1
    A = buf[:, :m]
    h = qr.AnyQR(m, n)
    s = _stream(torch)
    h.factor(buf.data_ptr(), ld, s)
    k = min(m, n)
    R = torch.full((n, k), 5.0, device="cuda")
    h.get_r(R.data_ptr(), k, s)
    torch.cuda.synchronize()
    assert torch.equal(buf[:, :m], A)                                # input untouched
    Rm = R.t().double()                                              # (k, n)
    assert torch.count_nonzero(torch.tril(Rm, -1)) == 0
    R64 = torch.linalg.qr(A.t().double(), mode="r").R
    sg = torch.sign(torch.diagonal(R64)) * torch.sign(torch.diagonal(Rm))
    assert (torch.max(torch.abs(Rm - sg[:, None] * R64)) / torch.max(torch.abs(R64))).item() <= 1e-4
    # thin Q = Q [I_k; 0] through the front-end's apply
    Qb = torch.zeros((k, m), device="cuda")
    Qb[torch.arange(k), torch.arange(k)] = 1.0
    h.apply_q(Qb.data_ptr(), m, k, False, s)
    torch.cuda.synchronize()
    Qm = Qb.t().double()                                             # (m, k)
    orth = (torch.linalg.norm(Qm.t() @ Qm - torch.eye(k, device="cuda", dtype=torch.float64)) / (max(n, 32) * EPS)).item()
    assert orth <= 5.0, orth
    back = (torch.linalg.norm(Qm @ Rm - A.t().double()) / (torch.linalg.norm(A.double()) * max(n, 32) * EPS)).item()
    assert back <= 5.0, back
    # Q^T (Q X) = X on odd column counts
    X = torch.rand((7, m), generator=g, device="cuda")
    Y = X.clone()
    h.apply_q(Y.data_ptr(), m, 7, False, s)
    h.apply_q(Y.data_ptr(), m, 7, True, s)
    torch.cuda.synchronize()
    assert (torch.linalg.norm(Y - X) / torch.linalg.norm(X)).item() <= 50 * EPS
    if m >= n:
        B = torch.rand((5, m), generator=g, device="cuda") * 2 - 1
        Xs = torch.zeros((5, n), device="cuda")
        h.solve_ls(B.data_ptr(), m, 5, Xs.data_ptr(), n, s)
        torch.cuda.synchronize()
        want = torch.linalg.lstsq(A.t().double(), B.t().double()).solution
        cond = torch.linalg.cond(A.t().double()).item()
        rel = (torch.linalg.norm(Xs.t().double() - want) / torch.linalg.norm(want)).item()
        assert rel <= max(1e-4, 20 * cond * EPS), (rel, cond)
    h.close()


def test_arbitrary_size_equals_padded_reference_layout(cuda, qr, po):
    """The padded buffer the front-end factors is the oracle's factorisation of the zero-padded matrix, element for
    element (same tolerance as the whole-matrix parity tests): the drop-in layout is unchanged, only embedded."""
    torch = cuda
    m, n = 200, 150
    rng = np.random.default_rng(9)
    a = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    h = qr.AnyQR(m, n)
    d = torch.from_numpy(a).cuda()
    h.factor(d.data_ptr(), m, _stream(torch))
    torch.cuda.synchronize()
    L = qr.lib()
    mp, np_ = L.b200qr_any_padded_rows(h._h), L.b200qr_any_padded_cols(h._h)
    assert (mp, np_) == (224, 160)
    out = np.zeros((np_, mp), dtype=np.float32)
    qr.check(L.b200qr_copy(out.ctypes.data_as(C.c_void_p), C.c_void_p(L.b200qr_any_factored(h._h)), out.nbytes))
    pad = np.zeros((np_, mp), dtype=np.float32)
    pad[:n, :m] = a
    ref_out, _ = po.tiled_qr(pad, mp, np_)
    assert np.max(np.abs(out - ref_out)) <= 1e-4 * max(1.0, float(np.max(np.abs(ref_out))))
    assert not out[n:, :].any() or np.max(np.abs(out[n:, :])) == 0.0        # padded columns stay zero
    h.close()
