"""The algebra the tensor-core sweep engine relies on, checked in NumPy on the CPU.

The reference applies a block's reflectors one after another (qrdecomp.c:620-644 SLARFT, 736-762 SSSRFT); the engine
applies them as  C <- C - V (T^T (V^T C))  with  T = (diag(1/tau) + striu(V^T V))^-1, and runs every product as 3xTF32
(hi = tf32(x) by integer round-half-up of the low 13 mantissa bits, lo = x - hi; hi*hi + hi*lo + lo*hi).  These tests pin
both identities independently of any GPU: the compact-WY form equals the sequential application, for the GEQRT block
(unit lower triangular V) and the TSQRT block (V = [I; V2]); the split is exact and the dropped lo*lo term is below 2^-22.
"""
import numpy as np
import pytest

B = 32


def sequential(V, tau, C):
    """H_{b-1} ... H_1 H_0 C with H_r = I - tau_r v_r v_r^T (reflector 0 first, like the reference's k loop)."""
    C = C.copy()
    for r in range(V.shape[1]):
        v = V[:, r]
        C -= tau[r] * np.outer(v, v @ C)
    return C


def compact_wy(V, tau, C):
    G = V.T @ V
    T = np.linalg.inv(np.diag(1.0 / tau) + np.triu(G, 1))
    return C - V @ (T.T @ (V.T @ C))


@pytest.mark.parametrize("kind", ["geqrt", "tsqrt"])
def test_compact_wy_equals_sequential_reflectors(kind):
    rng = np.random.default_rng(7)
    if kind == "geqrt":                         # unit lower triangular, b x b (tile (k,k), qrdecomp.c:522-523)
        V = np.tril(rng.uniform(-1, 1, (B, B)), -1) + np.eye(B)
    else:                                       # [I; V2], 2b x b (tile (l,k) below the diagonal, qrdecomp.c:686)
        V = np.vstack([np.eye(B), rng.uniform(-0.5, 0.5, (B, B))])
    tau = 2.0 / np.sum(V * V, axis=0)           # tau_r = 2 / (v_r^T v_r) = 2 / (1 + |v_ess|^2)
    C = rng.uniform(-1, 1, (V.shape[0], 4 * B))
    want = sequential(V, tau, C)
    got = compact_wy(V, tau, C)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    # T is upper triangular with tau on its diagonal, T[0:r, r] = -tau_r T[0:r, 0:r] (V[:, 0:r]^T v_r)  
    T = np.linalg.inv(np.diag(1.0 / tau) + np.triu(V.T @ V, 1))
    assert np.allclose(np.tril(T, -1), 0) and np.allclose(np.diag(T), tau)
    for r in range(1, B):
        assert np.allclose(T[:r, r], -tau[r] * T[:r, :r] @ (V[:, :r].T @ V[:, r]))


def tf32_hi(x):
    """The engine's split: add half an ulp of the 10-bit mantissa and clear the low 13 bits (tf32_round, qr_device.cuh)."""
    u = x.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_3xtf32_split_is_exact_and_tight():
    rng = np.random.default_rng(11)
    x = (rng.uniform(-1, 1, 200000) * 10.0 ** rng.integers(-6, 6, 200000)).astype(np.float32)
    hi = tf32_hi(x)
    lo = x - hi                                  # float32 subtraction
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0)          # hi has a 10-bit mantissa
    assert np.all(hi.astype(np.float64) + lo.astype(np.float64) == x.astype(np.float64))   # the split loses nothing
    assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -11 * (1 + 2.0 ** -10))                   # |lo| <= half a tf32 ulp
    # the low part fits the tensor core's tf32 operand to within 2^-11 of ITSELF, i.e. 2^-22 of x
    lo_t = tf32_hi(lo)
    y = (rng.uniform(-1, 1, x.shape)).astype(np.float32)
    yh = tf32_hi(y)
    yl = tf32_hi(y - yh)
    exact = x.astype(np.float64) * y.astype(np.float64)
    three = hi.astype(np.float64) * yh + hi.astype(np.float64) * yl + lo_t.astype(np.float64) * yh
    rel = np.abs(three - exact) / np.maximum(np.abs(exact), 1e-300)
    assert np.max(rel) <= 2.0 ** -20                                       # lo*lo dropped, lo rounded to tf32
