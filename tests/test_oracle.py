"""The oracle (oracle/tiledqr_oracle.c) against the reference's golden vectors.

tests/golden/* were produced by oracle/make_golden.py from the UNMODIFIED reference CPU
path (qrdecomp.c taskQRP_threads + gridscheduler.c) on the reference's own input recipe
(srand(5) + initMatrix RANDZO, qrdecomp.c:81,89).  The restatement must reproduce them bit
for bit; only then is it trusted as the checker for the CUDA path.
"""
import hashlib

import numpy as np
import pytest

SHAPES = [(32, 32), (64, 64), (128, 128), (96, 160), (512, 128), (256, 256), (1024, 1024)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("m,n", SHAPES)
def test_input_generator_matches_reference(po, golden, m, n):
    a = po.randzo(m, n)
    assert sha(a) == golden["factorisations"]["%dx%d" % (m, n)]["sha256_in"]


@pytest.mark.parametrize("m,n", SHAPES)
@pytest.mark.parametrize("order", ["static", "sched"])
def test_factorisation_bit_exact(po, golden, m, n, order):
    if order == "sched" and m * n > 512 * 512:
        pytest.skip("scheduler-driven order covered at the smaller shapes")
    rec = golden["factorisations"]["%dx%d" % (m, n)]
    a = po.randzo(m, n)
    out, tau = po.tiled_qr(a, m, n, order=order)
    assert sha(out) == rec["sha256_out"]
    assert sha(tau) == rec["sha256_tau"]
    assert int(np.count_nonzero(tau)) == rec["nnz_tau"]


def test_known_answers_from_survey(po, golden):
    # SURVEY 8c known answers, 128 x 128
    a = po.randzo(128, 128)
    out, tau = po.tiled_qr(a, 128, 128)
    A, T = out.T, tau.T
    s = golden["factorisations"]["128x128"]["samples"]
    assert a[0, 0] == np.float32(-0.06) and a[0, 1] == np.float32(0.04)
    assert A[0, 0] == np.float32(s["R00"]) and abs(A[0, 0] - (-6.763675)) < 1e-5
    assert A[0, 1] == np.float32(s["R01"]) and A[1, 1] == np.float32(s["R11"])
    assert A[1, 0] == np.float32(s["V10"]) and abs(A[1, 0] - (-0.01130738)) < 1e-7
    assert T[0, 0] == np.float32(s["tau0"]) and T[32, 0] == np.float32(s["tau_tile10_0"])


def test_small_outputs_elementwise(po, golden_small):
    for key in ("32x32", "64x64", "128x128", "96x160"):
        m, n = map(int, key.split("x"))
        a = golden_small["in_" + key]
        out, tau = po.tiled_qr(a, m, n)
        assert np.array_equal(out, golden_small["out_" + key])
        assert np.array_equal(tau, golden_small["tau_" + key])


def test_against_live_reference_if_built(po):
    """In the build container oracle/_ref/libref.so exists: compare directly as well."""
    if po.ref_lib() is None:
        pytest.skip("oracle/_ref not built on this machine")
    for (m, n) in [(64, 64), (160, 96), (256, 128)]:
        a = po.randzo(m, n, seed=11)
        ro, rt = po.ref_tiled_qr(a, m, n)
        oo, ot = po.tiled_qr(a, m, n)
        assert np.array_equal(ro, oo) and np.array_equal(rt, ot)


@pytest.mark.parametrize("key", ["1x1", "1x3", "3x1", "2x2", "3x3", "4x4", "7x3", "3x7", "16x4", "40x8", "32x32", "64x64"])
def test_scheduler_trace_matches_reference(po, golden, key):
    M, N = map(int, key.split("x"))
    tr = po.sched_trace(M, N)
    rec = golden["sched_traces"][key]
    assert len(tr) == rec["tasks"]
    text = "".join("%d %d %d %d\n" % tuple(r) for r in tr)
    assert hashlib.sha256(text.encode()).hexdigest() == rec["sha256"]
    if key in golden["sched_traces_full"]:
        assert tr.tolist() == golden["sched_traces_full"][key]


def test_closed_form_order(po):
    # the serial order has a closed form; oracle_tiled_qr_static relies on it
    for (M, N) in [(1, 1), (4, 4), (7, 3), (3, 7), (12, 5)]:
        want = []
        for k in range(min(M, N)):
            want.append((0, k, k, k))
            want += [(2, i, k, k) for i in range(k + 1, M)]
            for j in range(N - 1, k, -1):
                want.append((1, k, j, k))
                want += [(3, i, j, k) for i in range(k + 1, M)]
        assert [tuple(r) for r in po.sched_trace(M, N).tolist()] == want


def test_oracle_quality(po):
    # the bounds the CUDA path is later held to are met by the oracle itself
    for (m, n) in [(128, 128), (256, 128)]:
        a = po.randzo(m, n)
        out, tau = po.tiled_qr(a, m, n)
        back, orth, _ = po.quality(a, out, tau, m, n)
        assert back < 1.0 and orth < 3.0
        tau2 = po.tau_from_v(out, m, n, 32)
        assert np.max(np.abs(tau2 - tau)) < 2e-6


@pytest.mark.parametrize("m,n,rb", [(1536, 64, 48), (512, 128, 32), (768, 32, 48), (256, 256, 64)])
def test_streaming_tsqr_oracle(po, m, n, rb):
    """The TSQRT restatement with an n-wide tile, chained over blocks of rows from an empty triangle (what the product's
    streaming TSQR engine runs per strip): R equals float64 LAPACK's up to row signs, and the reflector blocks it leaves
    behind reproduce [R; 0] when replayed on [0; A] in float64 with tau from the oracle."""
    rng = np.random.default_rng(m + n)
    a = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    V, tau, R = po.tsqr_stream(a, m, n, rb)
    Rm = R.T.astype(np.float64)
    assert np.count_nonzero(np.tril(Rm, -1)) == 0
    R64 = np.linalg.qr(a.T.astype(np.float64), mode="r")
    sg = np.sign(np.diag(R64)) * np.sign(np.diag(Rm))
    assert np.max(np.abs(Rm - sg[:, None] * R64)) <= 1e-4 * np.max(np.abs(R64))
    top = np.zeros((n, n))
    bot = a.T.astype(np.float64).copy()
    Vd = V.T.astype(np.float64)
    for bi in range(m // rb):
        rows = slice(bi * rb, (bi + 1) * rb)
        for c in range(n):
            v2 = Vd[rows, c]
            w = float(tau[bi, c]) * (top[c, :] + v2 @ bot[rows, :])
            top[c, :] -= w
            bot[rows, :] -= np.outer(v2, w)
    scale = np.max(np.abs(top))
    assert np.max(np.abs(bot)) <= 2e-5 * scale and np.max(np.abs(top - Rm)) <= 2e-5 * scale
