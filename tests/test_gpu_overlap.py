"""cudaQRTask with overlapped copies (gpucalc.cu:1614-1670 does copy-in, kernel, copy-out one after the other).

The overlapped host path starts the persistent kernel first, streams the matrix in by tile columns (progress words
-1 -> 0 as a column's data arrive) and copies finished tile columns back while the kernel is still running.  It must
return exactly what the plain path returns: every tile sees the same operation sequence whatever the schedule.
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_legacy(L, host_ptr, m, n, ldm, env):
    env = dict(env, QRB200_QUIET="1")                 # restored below: later tests check the legacy prints
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        L.cudaQRTask(C.c_void_p(host_ptr), m, n, ldm, 128)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("m,n,ldm,extra", [
    (256, 256, 256, {}),                      # one-role grid (variant 1)
    (1024, 512, 1024, {}),                    # tall
    (512, 1024, 512, {}),                     # wide: tile columns >= K are final after the last step
    (2048, 2048, 2080, {}),                   # role-split SIMT engine (variant 2), ldm > m on the host
    (2048, 2048, 2048, {"QRB200_TC": "3"}),   # tensor-core engine
])
def test_overlapped_equals_plain(cuda, qr, po, m, n, ldm, extra):
    torch = cuda
    L = qr.lib()
    rng = np.random.default_rng(m * 7 + n)
    a = np.zeros((n, ldm), dtype=np.float32)
    a[:, :m] = rng.uniform(-1, 1, size=(n, m)).astype(np.float32)
    L.b200qr_release_cache()
    plain = a.copy()
    run_legacy(L, plain.ctypes.data, m, n, ldm, dict(extra, QRB200_OVERLAP="0"))
    pinned = torch.from_numpy(a.copy()).pin_memory()
    for rep in range(2):                      # the second call reuses the cached plan: no stale progress words
        pinned.copy_(torch.from_numpy(a))
        run_legacy(L, pinned.data_ptr(), m, n, ldm, dict(extra, QRB200_OVERLAP="1"))
        got = pinned.numpy()
        assert np.array_equal(got[:, :m], plain[:, :m]), "rep %d: max diff %g" % (rep, np.max(np.abs(got[:, :m] - plain[:, :m])))
        if ldm > m:
            assert np.array_equal(got[:, m:], a[:, m:])             # padding rows untouched
    # pageable host memory works too when forced (the copies are staged by the runtime)
    pag = a.copy()
    run_legacy(L, pag.ctypes.data, m, n, ldm, dict(extra, QRB200_OVERLAP="1"))
    assert np.array_equal(pag[:, :m], plain[:, :m])
    L.b200qr_release_cache()


def test_overlap_is_the_default_for_large_pinned_matrices(cuda, qr, po):
    """4096^2 pinned (64 MB >= 32 MB): the default call takes the overlapped path; the result equals the device API's."""
    torch = cuda
    L = qr.lib()
    m = n = 4096
    g = torch.Generator(device="cuda").manual_seed(77)
    A = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = A.clone()
    plan = qr.Plan(m, n)
    plan.factor(d.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    plan.close()
    h = A.cpu().pin_memory()
    L.b200qr_release_cache()
    os.environ.pop("QRB200_OVERLAP", None)
    run_legacy(L, h.data_ptr(), m, n, m, {})
    assert torch.equal(h, d.cpu())
    L.b200qr_release_cache()
