"""torchrun script: the block-cyclic multi-GPU factorisation (engine picked by the library: tensor-core engine 3 for
large matrices) against the single-GPU result of the same engine on rank 0.  usage: torchrun ... tools/dist_tc3_check.py n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import gpu_tiled_qr_decomposition_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
stream = torch.cuda.current_stream().cuda_stream
rk = q.DistRank(n, n, rank, world)
blobs = [None] * world
dist.all_gather_object(blobs, rk.export())
rk.connect(blobs)
g = torch.Generator(device="cuda").manual_seed(5)
A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
best = 1e9
for it in range(3):
    rk.upload(A0.data_ptr()); rk.reset(stream); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rk.factor(stream); e1.record(); torch.cuda.synchronize(); dist.barrier()
    best = min(best, e0.elapsed_time(e1))
t = torch.tensor([best], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
d = torch.empty_like(A0); rk.download(d.data_ptr())
own = rk.owned_tile_cols()
cols = torch.tensor([c for j in own for c in range(j * 32, (j + 1) * 32)], device="cuda")
# assemble the full result on rank 0
full = torch.zeros_like(A0)
full[cols] = d[cols]
dist.all_reduce(full, op=dist.ReduceOp.SUM)
if rank == 0:
    print("world %d n %d block %d grid %d: %.2f ms  %.1f TFLOP/s" % (world, n, rk.block, rk.grid, t.item(), (4 / 3) * n ** 3 / t.item() / 1e9), flush=True)
    plan = q.Plan(n, n)
    ref = A0.clone(); plan.factor(ref.data_ptr(), n, stream); torch.cuda.synchronize()
    print("single-GPU variant", plan.variant, " max|multi - single| =", (full - ref).abs().max().item(), " identical:", bool(torch.equal(full, ref)), flush=True)
    plan.close()
dist.barrier()
# ---- the same factorisation from HOST data with the copies overlapped (b200qr_dist_factor_host, packed layout): every
# owned column must come back exactly as the device run left it
import time
hb = A0[cols].cpu().pin_memory()                       # this rank's tile columns, consecutive (row of the tensor = matrix column)
want = d[cols].cpu()
best_h = 1e9
for it in range(3):
    work = hb.clone().pin_memory()
    dist.barrier()
    t0 = time.perf_counter()
    rk.reset_host(); dist.barrier()
    kms = rk.factor_host(work.data_ptr(), n, packed=True)
    dist.barrier()
    best_h = min(best_h, (time.perf_counter() - t0) * 1e3)
same = torch.tensor([1 if torch.equal(work, want) else 0], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print("host path (copies overlapped): %.2f ms wall, kernel %.2f ms, %.1f TFLOP/s end to end;  host identical: %s"
          % (best_h, kms, (4 / 3) * n ** 3 / best_h / 1e9, bool(same.item())), flush=True)
dist.barrier()
rk.close()
dist.destroy_process_group()
