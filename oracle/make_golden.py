"""Generate tests/golden/* from the UNMODIFIED reference CPU path (oracle/_ref/libref.so).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle`); the GPU box and the test-suite read the committed fixtures.

    python oracle/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(po.HERE), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert po.ref_lib() is not None, "oracle/_ref/libref.so missing: run make -C oracle"
    os.makedirs(OUT, exist_ok=True)
    meta = {"generator": "oracle/make_golden.py", "source": "oracle/_ref/libref.so "
            "(qrdecomp.c taskQRP_threads, 8 pthreads, b=32, useWY=1; srand(5)+initMatrix RANDZO)",
            "factorisations": {}, "sched_traces": {}}
    # factorisations: full arrays for the small ones, digests + samples for all
    full = {}
    for (m, n) in [(32, 32), (64, 64), (128, 128), (96, 160), (512, 128), (256, 256), (1024, 1024)]:
        a = po.randzo(m, n)
        out, tau = po.ref_tiled_qr(a, m, n)
        A = out.T
        T = tau.T
        rec = {"sha256_in": sha(a), "sha256_out": sha(out), "sha256_tau": sha(tau),
               "nnz_tau": int(np.count_nonzero(tau)),
               "samples": {"A00_in": float(a[0, 0]), "A10_in": float(a[0, 1]), "A01_in": float(a[1, 0]),
                           "R00": float(A[0, 0]), "R01": float(A[0, 1]),
                           "R11": float(A[1, 1]) if min(m, n) > 1 else None,
                           "V10": float(A[1, 0]), "tau0": float(T[0, 0]), "tau1": float(T[1, 0]),
                           "tau_tile10_0": float(T[32, 0]) if m > 32 else None}}
        meta["factorisations"]["%dx%d" % (m, n)] = rec
        if m * n <= 160 * 160:
            full["in_%dx%d" % (m, n)] = a
            full["out_%dx%d" % (m, n)] = out
            full["tau_%dx%d" % (m, n)] = tau
    np.savez_compressed(os.path.join(OUT, "ref_small.npz"), **full)
    # scheduler traces (serial single-worker order of gridscheduler.c)
    small = {}
    for (M, N) in [(1, 1), (1, 3), (3, 1), (2, 2), (3, 3), (4, 4), (7, 3), (3, 7), (16, 4), (40, 8), (32, 32), (64, 64)]:
        tr = po.ref_sched_trace(M, N)
        text = "".join("%d %d %d %d\n" % tuple(r) for r in tr)
        meta["sched_traces"]["%dx%d" % (M, N)] = {"tasks": int(len(tr)),
                                                   "sha256": hashlib.sha256(text.encode()).hexdigest()}
        if len(tr) <= 200:
            small["%dx%d" % (M, N)] = tr.tolist()
    meta["sched_traces_full"] = small
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
