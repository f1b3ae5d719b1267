"""torchrun script: times the block-cyclic multi-GPU factorisation under several plan settings in ONE process group
(QRB200_LAG / QRB200_NP are read when the plan is created).  usage: torchrun ... tools/dist_scan.py n "LAG:NP" "LAG:NP" ...
("-" = the library default).  Rank 0 prints one line per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1])
settings = sys.argv[2:] or ["-:-"]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(5)
A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
for st in settings:
    lag, npc = st.split(":")
    for k, v in (("QRB200_LAG", lag), ("QRB200_NP", npc)):
        if v == "-":
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    rk = q.DistRank(n, n, rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, rk.export())
    rk.connect(blobs)
    times = []
    for it in range(3):
        rk.upload(A0.data_ptr()); rk.reset(stream); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rk.factor(stream); e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(t.item())
    if rank == 0:
        best = min(times[1:])
        print("world %d n %d LAG %s NP %s variant %d grid %d: %.2f ms  %.1f TFLOP/s  (runs: %s)" % (
            world, n, lag, npc, rk.variant, rk.grid, best, (4 / 3) * n ** 3 / best / 1e9, " ".join("%.1f" % x for x in times)), flush=True)
    dist.barrier()
    rk.close()
    torch.cuda.synchronize()
    dist.barrier()
dist.destroy_process_group()
