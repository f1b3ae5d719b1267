"""Tensor-core sweep engine (csrc/qr_persistent_tc3*, b200qr_opts.engine = 2): same factorisation, the trailing update
runs as 3xTF32 tcgen05 products.  Stated tolerance: elementwise within 5e-4 * max|R| of the oracle (3xTF32 keeps
~2^-21 per product instead of 2^-24), backward error and orthogonality <= 10 in units of n*eps (SURVEY 8c's bound for
the 3xTF32 path)."""
import os

import numpy as np
import pytest

from qr_checks import probe_accuracy

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float32).eps)


@pytest.mark.parametrize("m,n", [(4096, 3072), (4096, 4096)])
def test_tc_sweeps_match_oracle_up_to_row_sign(cuda, qr, po, m, n):
    """Past ~3000 pivots the elementwise comparison stops being meaningful for a 3xTF32 engine: a GEQRT pivot x0 that is
    within the engine's rounding level of zero (|x0| < ~1e-3 here, six to ten of 3072 pivots) gets the other sign
    (qrdecomp.c:1208 sign(x0)), which flips that row of R and changes the reflectors below it.  BASELINE.json's criterion
    is exactly this: R equal to the reference's up to row sign, plus backward error and orthogonality."""
    from qr_checks import r_rel_err_up_to_row_sign
    torch = cuda
    a = po.randzo(m, n)
    d = torch.from_numpy(a).cuda()
    plan = qr.Plan(m, n, engine=2)
    plan.factor(d.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref_out, _ = po.tiled_qr(a, m, n)
    err, flips = r_rel_err_up_to_row_sign(d.cpu().numpy(), ref_out, m, n)
    assert err <= 5e-4 and flips <= 16, (err, flips)
    back, orth = probe_accuracy(torch, plan, d, torch.from_numpy(a).cuda(), m, n, 64)
    assert orth <= 10.0 and back <= 10.0, (back, orth)
    plan.close()


@pytest.mark.parametrize("m,n", [(2048, 2048), (3072, 2048), (2048, 2560)])
def test_tc_sweeps_match_oracle(cuda, qr, po, m, n):
    torch = cuda
    a = po.randzo(m, n)
    d = torch.from_numpy(a).cuda()
    dt = torch.zeros_like(d)
    plan = qr.Plan(m, n, engine=2)
    assert plan.variant == 5
    s = torch.cuda.current_stream().cuda_stream
    plan.factor(d.data_ptr(), m, s)
    plan.tau_to_reference(dt.data_ptr(), m, s)
    torch.cuda.synchronize()
    out = d.cpu().numpy()
    ref_out, ref_tau = po.tiled_qr(a, m, n)
    scale = float(np.max(np.abs(np.triu(ref_out.T))))
    err = float(np.max(np.abs(out - ref_out)))
    assert err <= 5e-4 * scale, (err, scale)
    assert err <= 1e-2                                   # far inside the reference's own 1e-3-per-1 criterion scaled by |R|
    back, orth = probe_accuracy(torch, plan, d, torch.from_numpy(a).cuda(), m, n, 64)
    assert orth <= 10.0 and back <= 10.0, (back, orth)
    plan.close()


def _factor(torch, qr, A, n, engine):
    plan = qr.Plan(n, n, engine=engine)
    d = A.clone()
    plan.factor(d.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return plan, d


def test_tc3_large_matrix_against_simt_engine(cuda, qr):
    """Engine 3 (the default above 12288^2) at a size the oracle does not finish in seconds: R must equal the SIMT
    engine's up to row sign (BASELINE.json's criterion; a pivot close to zero may round to the other sign), within
    5e-4 * max|R|, and pass the backward-error / orthogonality bounds (<= 10 n eps) through GPU-applied Q."""
    torch = cuda
    n = 6144
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    p_simt, d_simt = _factor(torch, qr, A, n, 1)
    p_tc, d_tc = _factor(torch, qr, A, n, 2)
    assert p_simt.variant == 2 and p_tc.variant == 5
    Rs, Rt = torch.triu(d_simt.t()), torch.triu(d_tc.t())
    sg = torch.sign((Rs * Rt).sum(dim=1))
    scale = Rs.abs().max().item()
    assert int((sg == 0).sum().item()) == 0
    assert ((Rt - sg[:, None] * Rs).abs().max() / scale).item() <= 5e-4
    back, orth = probe_accuracy(torch, p_tc, d_tc, A, n, n, 32)
    assert orth <= 10.0 and back <= 10.0, (back, orth)
    p_simt.close(); p_tc.close()


def test_default_engine_selection(cuda, qr):
    """Automatic choice (opts.engine = 0, QRB200_TC unset): a function of the shape only -- SIMT engine up to 8192^2,
    tensor-core engine from 12288^2 on one GPU; the choice is reported by b200qr_plan_variant."""
    old = os.environ.pop("QRB200_TC", None)
    try:
        for n, want in ((8192, 2), (12288, 5)):
            plan = qr.Plan(n, n)
            assert plan.variant == want, (n, plan.variant)
            plan.close()
    finally:
        if old is not None:
            os.environ["QRB200_TC"] = old


@pytest.mark.parametrize("n", [16384, 32768])
def test_default_engine_at_headline_sizes(cuda, qr, n):
    """The engine the headline numbers run on, at the sizes it is the default for (BASELINE.json's 32768^2 target and
    half of it), checked in fp64 through GPU-applied Q: backward error and orthogonality <= 10 n eps, R^T R = A^T A on
    the leading 2048 columns, every diagonal entry of R non-zero.  The oracle cannot reach these sizes."""
    torch = cuda
    free, _ = torch.cuda.mem_get_info()
    need = 4 * n * n * (2 + 6) + (3 << 30)
    if free < need:
        pytest.skip("needs %.0f GB of free device memory" % (need / 2 ** 30))
    g = torch.Generator(device="cuda").manual_seed(6)
    A = torch.rand((n, n), device="cuda", generator=g) * 2 - 1
    plan, d = _factor(torch, qr, A, n, 0)
    assert plan.variant == 5
    back, orth = probe_accuracy(torch, plan, d, A, n, n, 32)
    assert orth <= 10.0 and back <= 10.0, (back, orth)
    c1 = 2048
    R = torch.triu(d[:c1, :c1].t().double())
    Ad = A[:c1].t().double()
    G2 = Ad.t() @ Ad
    gram = (torch.linalg.norm(R.t() @ R - G2) / torch.linalg.norm(G2)).item()
    assert gram <= 0.1 * n * EPS, gram              # same unit as the backward-error bound (measured: 0.02 n eps)
    assert float(torch.diagonal(d).abs().min().item()) > 0.0
    plan.close()
