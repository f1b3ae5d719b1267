"""apply-Q and the accuracy bounds it makes checkable at any size:
backward error |A - QR| and orthogonality |Q^T Q - I| with Q applied by the GPU kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float32).eps)


def factor(qr, torch, A):
    n, m = A.shape
    d = A.clone()
    plan = qr.Plan(m, n)
    plan.factor(d.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return plan, d


@pytest.mark.parametrize("m,n", [(256, 256), (512, 256), (1024, 1024)])
def test_qt_a_is_r_and_q_r_is_a(cuda, qr, po, m, n):
    torch = cuda
    A = torch.from_numpy(po.randzo(m, n)).cuda()
    plan, d = factor(qr, torch, A)
    s = torch.cuda.current_stream().cuda_stream
    # Q^T A = [R; 0]
    C = A.clone()
    plan.apply_q(d.data_ptr(), m, C.data_ptr(), m, n, True, s)
    torch.cuda.synchronize()
    R = torch.triu(d.t())                                   # (m, n)
    err = torch.linalg.norm(C.t().double() - R.double()) / (torch.linalg.norm(A.double()) * n * EPS)
    assert err.item() <= 5.0, err.item()
    # Q [R; 0] = A  (backward error of the factorisation)
    C = R.t().contiguous()
    plan.apply_q(d.data_ptr(), m, C.data_ptr(), m, n, False, s)
    torch.cuda.synchronize()
    back = torch.linalg.norm(C.double() - A.double()) / (torch.linalg.norm(A.double()) * n * EPS)
    assert back.item() <= 5.0, back.item()
    # against the oracle's float64 replay of the same reflectors (pins the operation order)
    if m <= 512:
        out = d.cpu().numpy()
        tau = np.zeros_like(out)
        dt = torch.zeros_like(d)
        plan.tau_to_reference(dt.data_ptr(), m, s)
        torch.cuda.synchronize()
        I = torch.eye(m, device="cuda").contiguous()
        Cq = I[:, :].clone()                                 # m x m identity, column-major == itself
        plan.apply_q(d.data_ptr(), m, Cq.data_ptr(), m, m, True, s)
        torch.cuda.synchronize()
        want = po.apply_qt_f64(out, dt.cpu().numpy(), m, n, 32, np.eye(m))
        assert np.max(np.abs(Cq.t().cpu().numpy() - want)) <= 2e-5
    plan.close()


def test_orthogonality_and_backward_error_at_4096(cuda, qr):
    """Full-size accuracy through random probes: |Q^T (Q X) - X| and |A X - Q (R X)|."""
    torch = cuda
    m = n = 4096
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    plan, d = factor(qr, torch, A)
    s = torch.cuda.current_stream().cuda_stream
    X = torch.randn((64, m), generator=g, device="cuda")       # 64 probe columns (column-major m x 64)
    Y = X.clone()
    plan.apply_q(d.data_ptr(), m, Y.data_ptr(), m, 64, False, s)
    plan.apply_q(d.data_ptr(), m, Y.data_ptr(), m, 64, True, s)
    torch.cuda.synchronize()
    orth = (torch.linalg.norm(Y.double() - X.double()) / (torch.linalg.norm(X.double()) * n * EPS)).item()
    assert orth <= 5.0, orth
    R = torch.triu(d.t())                                       # (m, n)
    RX = (R.double() @ X.t().double()).float().t().contiguous()  # column-major m x 64
    plan.apply_q(d.data_ptr(), m, RX.data_ptr(), m, 64, False, s)
    torch.cuda.synchronize()
    AX = A.t().double() @ X.t().double()
    back = (torch.linalg.norm(RX.t().double() - AX) / (torch.linalg.norm(A.double()) * torch.linalg.norm(X.double()) / np.sqrt(64) * n * EPS)).item()
    assert back <= 5.0, back
    plan.close()
