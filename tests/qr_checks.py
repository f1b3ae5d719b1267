"""Accuracy of a factorisation through GPU-applied Q on random probe vectors (the way to check
sizes the oracle cannot reach).  Both numbers are in units of n * eps(float32), the unit BASELINE.json's criterion uses:
    orthogonality  |Q^T (Q X) - X|_F / (|X|_F n eps)
    backward error |A Xn - Q (R Xn)|_F / (|A|_F |Xn|_F / sqrt(ncols) n eps)
fp64 is used for the products with A and R only; Q is applied by b200qr_apply_q (float32, the stored reflectors)."""
import numpy as np

EPS = float(np.finfo(np.float32).eps)


def probe_accuracy(torch, plan, d, A0, m, n, ncols=32, seed=3):
    """plan: a qr.Plan that has just factored A0 (storage (n, m) = column-major m x n) into d."""
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn((ncols, m), generator=g, device="cuda")
    Y = X.clone()
    plan.apply_q(d.data_ptr(), m, Y.data_ptr(), m, ncols, False, s)
    plan.apply_q(d.data_ptr(), m, Y.data_ptr(), m, ncols, True, s)
    torch.cuda.synchronize()
    orth = (torch.linalg.norm(Y.double() - X.double()) / (torch.linalg.norm(X.double()) * n * EPS)).item()
    del Y
    Xn = torch.randn((ncols, n), generator=g, device="cuda")
    K = min(m, n)
    RX = torch.zeros((ncols, m), device="cuda")
    # R Xn, R = upper triangle of the first K rows: one tile row of columns at a time keeps the fp64 temporaries small
    step = 4096
    for r0 in range(0, K, step):
        r1 = min(K, r0 + step)
        Rblk = torch.triu(d[:, r0:r1].t().double(), diagonal=r0)          # rows r0..r1 of R, (r1-r0) x n
        RX[:, r0:r1] = (Rblk @ Xn.t().double()).t().float()
        del Rblk
    plan.apply_q(d.data_ptr(), m, RX.data_ptr(), m, ncols, False, s)
    torch.cuda.synchronize()
    AX = torch.zeros((m, ncols), device="cuda", dtype=torch.float64)
    anorm2 = 0.0
    for c0 in range(0, n, step):
        c1 = min(n, c0 + step)
        Ablk = A0[c0:c1].double()                                          # columns c0..c1 of A, stored (cols, m)
        AX += Ablk.t() @ Xn[:, c0:c1].t().double()
        anorm2 += float((Ablk * Ablk).sum().item())
        del Ablk
    back = (torch.linalg.norm(RX.t().double() - AX) /
            (anorm2 ** 0.5 * torch.linalg.norm(Xn.double()) / ncols ** 0.5 * n * EPS)).item()
    return back, orth


def r_rel_err_up_to_row_sign(out, ref_out, m, n):
    """BASELINE.json's criterion for R: equal to the reference's up to the sign of each row.  out / ref_out are factored
    matrices in (n, m) storage (column-major m x n); returns (max |R - s R_ref| / max |R_ref|, number of flipped rows).
    A row's sign is decided by a pivot x0 in GEQRT (qrdecomp.c:1208): when |x0| is at the rounding level of the engine
    (3xTF32 products leave ~1e-5 max|R|), two correct implementations may pick opposite signs for that row."""
    K = min(m, n)
    R = np.triu(out.T[:K, :].astype(np.float64))
    Rr = np.triu(ref_out.T[:K, :].astype(np.float64))
    s = np.sign(np.sum(R * Rr, axis=1))
    s[s == 0] = 1.0
    err = float(np.max(np.abs(R - s[:, None] * Rr)) / np.max(np.abs(Rr)))
    return err, int(np.sum(s < 0))
