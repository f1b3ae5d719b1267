"""torchrun script: peer-publish stress test of the multi-GPU tensor-core engine.  The same
n x n matrix is factored `iters` times on W GPUs; every rank compares its owned tile columns bit for bit with its first
run (a stale or early-seen operand tile would change them), and rank 0 compares the first assembled result with the
single-GPU engine's.  usage: torchrun ... tools/dist_stress.py n iters [engine: 2 tensor-core (default), 1 SIMT]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
engine = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.current_stream().cuda_stream
rk = q.DistRank(n, n, rank, world, engine=engine)
blobs = [None] * world
dist.all_gather_object(blobs, rk.export())
rk.connect(blobs)
g = torch.Generator(device="cuda").manual_seed(5)
A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
own = rk.owned_tile_cols()
mask = torch.zeros(n // 32, device="cuda", dtype=torch.bool)
mask[torch.tensor(own, device="cuda")] = True
first = None
bad = 0
d = torch.empty_like(A0)
for it in range(iters):
    rk.upload(A0.data_ptr()); rk.reset(stream); dist.barrier()
    rk.factor(stream); torch.cuda.synchronize()
    dist.barrier()          # peers read this rank's tiles / write into its buffers until THEIR kernels are done (include/qrb200.h)
    rk.download(d.data_ptr())
    mine = d.view(n // 32, 32 * n)[mask]
    if first is None:
        first = mine.clone()
    elif not torch.equal(first, mine):
        bad += 1
t = torch.tensor([bad], device="cuda"); dist.all_reduce(t)
d.view(n // 32, 32 * n).mul_(mask[:, None].float())
dist.reduce(d, dst=0)
if rank == 0:
    plan = q.Plan(n, n, engine=engine)
    ref = A0.clone(); plan.factor(ref.data_ptr(), n, stream); torch.cuda.synchronize()
    print("stress: world %d n %d variant %d, %d factorisations: %d differ from the first; last assembled result identical to the "
          "single-GPU engine: %s" % (world, n, rk.variant, iters, int(t.item()), bool(torch.equal(ref, d))), flush=True)
    plan.close()
dist.barrier()
rk.close()
dist.destroy_process_group()
