"""N > 1 host logic on CPU: world_size 2 over gloo (see tests/dist_worker.py)."""
import os
import socket
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("m,n,G", [(256, 256, 2), (384, 256, 3), (256, 384, 4), (512, 512, 8)])
def test_block_cyclic_two_ranks_gloo(qr, po, tmp_path, m, n, G):
    import torch.multiprocessing as mp
    import dist_worker
    res = str(tmp_path / "result.txt")
    mp.spawn(dist_worker.run, args=(2, free_port(), m, n, G, res), nprocs=2, join=True)
    assert open(res).read() == "OK"


def test_rank_tickets_partition(qr):
    """Ownership maps for 1..8 ranks: disjoint, complete, block-cyclic."""
    import ctypes as C
    import numpy as np
    L = qr.lib()
    for (M, N, G) in [(20, 20, 4), (64, 64, 16), (40, 24, 8), (24, 40, 8)]:
        total = L.b200qr_sched_sweep_order(M, N, G, None, 0)
        for world in (1, 2, 3, 4, 8):
            cnt, pan = 0, []
            for r in range(world):
                sw, pa = qr.rank_tickets(M, N, G, r, world)
                assert all(g % world == r for _, g in sw.tolist())
                cnt += len(sw)
                pan += pa.tolist()
            assert cnt == total and sorted(pan) == list(range(min(M, N)))


@pytest.mark.parametrize("m_local,n,world", [(512, 128, 2), (256, 64, 3)])
def test_tsqr_exchange_gloo(qr, po, tmp_path, m_local, n, world):
    """configs[3] across ranks: local reduction, all-gather of the R factors, redundant top reduction (TsqrDist)."""
    import torch.multiprocessing as mp
    import dist_worker
    res = str(tmp_path / "result.txt")
    mp.spawn(dist_worker.run_tsqr, args=(world, free_port(), m_local, n, res), nprocs=world, join=True)
    assert open(res).read() == "OK"
