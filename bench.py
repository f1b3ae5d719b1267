#!/usr/bin/env python
"""bench.py -- QR GFLOP/s (flops = 2mn^2 - 2n^3/3) of the B200 tiled QR, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size n]

A step = one tiled QR factorisation of a synthetic n x n float32 matrix (default 8192 x 8192,
BASELINE.json configs[1]) with the reference's tile size 32.  `value` is measured with the
matrix resident in HBM (CUDA events on the launching stream around the factorisation only,
like the reference's own "GPU: x ms" line, gpucalc.cu:1635-1647); `e2e` is the same metric
through the legacy host entry point cudaQRTask() on a pinned HOST buffer, H2D and D2H copies
inside the timed region.  Between timed steps the 256 MiB input is restored from a pristine
copy, which also evicts the 126 MB L2.  `--impl reference` times the reference's own CPU path
(oracle/_ref, 8 pthreads) on a bounded sample.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "QR GFLOP/s (2mn^2-2n^3/3)"


def qr_flops(m, n):
    return 2.0 * m * n * n - 2.0 * n ** 3 / 3.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(n, reps=1):
    """The reference CPU path (taskQRP_threads, 8 pthreads, qrdecomp.c:145) on an n x n RANDZO
    matrix; returns (best seconds, kind, cores)."""
    import numpy as np
    import pyoracle as po
    po.build()
    a = po.randzo(n, n)
    best = 1e30
    ref = po.ref_lib(fast=True)
    for _ in range(reps):
        if ref is not None:
            out, tau = np.zeros_like(a), np.zeros_like(a)
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            sys.stdout.flush()
            os.dup2(devnull, 1)                       # the reference prints "CPU: x ms" from C
            try:
                t0 = time.perf_counter()
                ref.taskQRP_threads(a, out, tau, n, n, 32, n, 1)
                dt = time.perf_counter() - t0
            finally:
                os.dup2(saved, 1); os.close(devnull); os.close(saved)
            kind, cores = "reference", 8
        else:
            out, tau = a.copy(), np.zeros_like(a)
            cores = min(8, os.cpu_count() or 1)
            t0 = time.perf_counter()
            po.lib(fast=True).oracle_tiled_qr_threads(out, tau, n, n, 32, n, cores)
            dt = time.perf_counter() - t0
            kind = "port"
        best = min(best, dt)
    return best, kind, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = args.ref_size or (4096 if args.steps + args.warmup <= 10 else 2048)
    for _ in range(args.warmup):
        cpu_reference_run(n)
    ts = [cpu_reference_run(n) for _ in range(args.steps)]
    sec = sum(t[0] for t in ts) / len(ts)
    kind, cores = ts[0][1], ts[0][2]
    val = qr_flops(n, n) / sec * 1e-9
    sample = "%dx%d RANDZO matrix, b=32, taskQRP_threads with %d pthreads, -O3 -march=x86-64-v3" % (n, n, cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%dx%d square tiled QR, b=32 (bounded sample of the %s workload of the b200 arm: the "
                                   "reference's O(p*q)-scan scheduler needs minutes at 8192 and has no multi-GPU path)"
                                   % (n, n, "8192x8192" if args.gpus <= 1 else "%dx%d" % (args.multi_size, args.multi_size))},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        return run_b200_multi(args, rank, world, local_rank)
    return run_b200_single(args)


def gram_check(torch, d, A0, c0, c1):
    """|R^T R - A^T A| / |A^T A| restricted to matrix columns [c0, c1): needs only those columns of
    the factored matrix d (R on/above the diagonal) and of the input A0.  Storage is (n, ld)."""
    rows = torch.arange(c1, device=d.device)[:, None]
    cols = torch.arange(c0, c1, device=d.device)[None, :]
    R = torch.where(rows <= cols, d[c0:c1, :c1].t().double(), torch.zeros((), dtype=torch.float64, device=d.device))
    Ad = A0[c0:c1].t().double()
    G2 = Ad.t() @ Ad
    return (torch.linalg.norm(R.t() @ R - G2) / torch.linalg.norm(G2)).item()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the `ncu --set full` captures summarised under profiles/
# (r1_simt_8192_summary.txt, r1_tc3_16384_summary.txt), keyed by (kernel variant, n); other shapes report null.
NCU_TRAFFIC_BYTES = {(2, 8192): 4.175e9 + 7.903e9, (5, 16384): 189.27e9 + 177.82e9}


def roofline_of(variant, achieved_tf, fp32_peak_tf, peaks, peak_src, n, ms_step):
    """The factorisation is one persistent kernel, so the kernel's roofline is the step's.  Variants 1 / 2 apply the
    reflectors on the FP32 SIMT pipe (bound: the FFMA rate, measured in this process).  Variant 5 runs the trailing
    update as 3xTF32 tcgen05 products: the bound is the tensor pipe at a sixth of the measured dense bf16 rate (TF32
    runs at half the bf16 rate, three passes per product); the fraction of the FP32 SIMT peak -- the north-star's
    '% of B200 peak for the working precision' -- is reported next to it."""
    common = {"achieved": achieved_tf, "unit": "TFLOP/s", "traffic": NCU_TRAFFIC_BYTES.get((variant, n)),
              "traffic_note": "DRAM bytes per launch (ncu), vs 8*m*n = %.3g algorithmic: the right-looking sweeps stream the "
                              "trailing tiles once per step; L2 absorbs what fits" % (8.0 * n * n),
              "frac_of_fp32_simt_peak": achieved_tf / fp32_peak_tf, "fp32_simt_peak": fp32_peak_tf,
              "frac_of_bf16_tensor_peak": achieved_tf / peaks["bf16_tflops"],
              "hbm_algorithmic_gbs": 8.0 * n * n / (ms_step * 1e-3) * 1e-9}
    if variant == 5:
        peak = peaks["bf16_tflops"] / 6.0
        common.update({"bound": "tensor", "peak": peak, "frac": achieved_tf / peak,
                       "kernel": "tiledqr_persistent_tc3 (the only kernel of a step)",
                       "peak_source": "%s bf16_tflops / 6 (3xTF32: TF32 at half the bf16 rate, three passes); FP32 SIMT "
                                      "peak from the FFMA microbench run in this process" % peak_src})
    else:
        common.update({"bound": "fp32-simt", "peak": fp32_peak_tf, "frac": achieved_tf / fp32_peak_tf,
                       "kernel": "tiledqr_persistent (the only kernel of a step)",
                       "peak_source": "FFMA microbench run in this process (b200qr_measure_fp32_peak); %s carries "
                                      "only HBM and bf16-tensor peaks" % peak_src})
    return common


def north_star_run(torch, q, n, fp32_peak_tf, peaks, peak_src):
    """One warm-up and two timed factorisations of an n x n matrix (BASELINE.json's 32768 target on one GPU)."""
    flops = qr_flops(n, n)
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(6)
    A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = torch.empty_like(A0)
    plan = q.Plan(n, n)
    best = None
    for it in range(3):
        d.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.factor(d.data_ptr(), n, stream)
        e1.record()
        torch.cuda.synchronize()
        if it > 0:
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
    tf = flops / (best * 1e-3) * 1e-12
    out = {"workload": "%dx%d square tiled QR, tile size 32, 1 GPU" % (n, n), "ms": best, "value": tf * 1e3, "unit": "GFLOP/s",
           "kernel_variant": plan.variant, "gram_rel_err": gram_check(torch, d, A0, 0, 1024),
           "roofline": roofline_of(plan.variant, tf, fp32_peak_tf, peaks, peak_src, n, best)}
    plan.close()
    del A0, d
    torch.cuda.empty_cache()
    return out


def run_b200_single(args):
    import torch
    import gpu_tiled_qr_decomposition_b200 as q
    q.build()
    L = q.lib()
    n = args.size
    flops = qr_flops(n, n)
    stream = torch.cuda.current_stream().cuda_stream

    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1   # uniform[-1,1)
    d = torch.empty_like(A0)
    plan = q.Plan(n, n)
    for _ in range(args.warmup):
        d.copy_(A0)
        plan.factor(d.data_ptr(), n, stream)
    torch.cuda.synchronize()

    sampler = ClockSampler(0)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for e0, e1 in evs:
        d.copy_(A0)                       # restore the input; the copy also evicts L2
        e0.record()
        plan.factor(d.data_ptr(), n, stream)
        e1.record()
    torch.cuda.synchronize()
    ms_step = sum(e0.elapsed_time(e1) for e0, e1 in evs) / args.steps
    clocks = sampler.stop()
    value = flops / (ms_step * 1e-3) * 1e-9
    gram = gram_check(torch, d, A0, 0, min(n, 2048))

    # ---- e2e: the legacy host entry point on a pinned host buffer
    os.environ["QRB200_QUIET"] = "1"
    h0 = A0.cpu()
    h = torch.empty((n, n), dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(2):                                            # warm-up (context, pinned pages)
        h.copy_(h0)
        L.cudaQRTask(C.c_void_p(h.data_ptr()), n, n, n, 128)
    e2e_s = 0.0
    for _ in range(e2e_steps):
        h.copy_(h0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.cudaQRTask(C.c_void_p(h.data_ptr()), n, n, n, 128)
        e2e_s += time.perf_counter() - t0
    e2e_val = flops / (e2e_s / e2e_steps) * 1e-9
    e2e_err = float((h - d.cpu()).abs().max())                     # whole matrix, vs the device-resident path

    peaks, peak_src = measured_peaks()
    fp32 = C.c_double(0)
    q.check(L.b200qr_measure_fp32_peak(C.byref(fp32), 20000))
    achieved_tf = flops / (ms_step * 1e-3) * 1e-12
    roofline = roofline_of(plan.variant, achieved_tf, fp32.value * 1e-3, peaks, peak_src, n, ms_step)
    # the north-star configuration (32768 x 32768 on one GPU), one extra factorisation: reported next to the headline
    north = None
    if args.north_size and args.north_size != n:
        try:
            north = north_star_run(torch, q, args.north_size, fp32.value * 1e-3, peaks, peak_src)
        except Exception as e:                                    # e.g. not enough free memory on a shared box
            north = {"skipped": str(e)[:200]}
    sec, kind, cores = cpu_reference_run(args.cpu_sample, reps=1)
    cpu = {"value": qr_flops(args.cpu_sample, args.cpu_sample) / sec * 1e-9, "unit": "GFLOP/s", "cores": cores,
           "kind": kind, "sample": "%dx%d RANDZO matrix, b=32, one run of taskQRP_threads (%.1f s), "
           "-O3 -march=x86-64-v3, host has %d cpus" % (args.cpu_sample, args.cpu_sample, sec, os.cpu_count())}
    line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform[-1,1), torch cuda generator seed 5",
            "config": {"workload": "%dx%d square tiled QR, tile size 32 (reference default), flat-tree TS kernels"
                                   % (n, n), "parallelism": "1 GPU",
                       "l2": "input restored from a pristine %d MiB copy before every timed step (> 126 MB L2)"
                             % (4 * n * n >> 20),
                       "persistent_grid_ctas": plan.grid, "kernel_variant": plan.variant, "tile_tasks": plan.tile_tasks,
                       "macro_tasks": plan.macro_tasks},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n,
                    "d2h_bytes_per_step": 4 * n * n, "api": "cudaQRTask(host float*, m, n, ldm, maxblocks) on pinned "
                    "host memory (workspace cached between calls of one shape; copies overlapped with the kernel: tile "
                    "columns stream in behind progress words, finished columns stream out)", "steps": e2e_steps,
                    "max_abs_diff_vs_device_path": e2e_err},
            "gpu_launches": args.steps * 1, "clocks": clocks, "check": {"gram_rel_err": gram}}
    if north is not None:
        line["north_star_32768"] = north
    print(json.dumps(line), flush=True)
    plan.close()


def run_b200_multi(args, rank, world, local_rank):
    """N > 1: one 32768 x 32768 matrix (BASELINE.json configs[4]), tile columns dealt to the ranks
    1-D block-cyclically; every rank runs the persistent kernel on its own GPU and reads the other
    ranks' reflector tiles over NVLink from inside it.  Strong scaling: total work is fixed."""
    import torch
    import torch.distributed as dist
    import gpu_tiled_qr_decomposition_b200 as q
    q.build()
    L = q.lib()
    n = args.multi_size
    flops = qr_flops(n, n)
    stream = torch.cuda.current_stream().cuda_stream

    rk = q.DistRank(n, n, rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, rk.export())
    rk.connect(blobs)
    g = torch.Generator(device="cuda").manual_seed(5)            # same matrix on every rank
    A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    own = rk.owned_tile_cols()

    def one_step(timed):
        rk.upload(A0.data_ptr())                                 # restore (device-to-device, evicts L2)
        rk.reset(stream)
        torch.cuda.synchronize()
        dist.barrier()                                           # every rank has reset its progress words
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rk.factor(stream)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    ms = sum(one_step(True) for _ in range(args.steps))
    dist.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = flops / (ms_step * 1e-3) * 1e-9

    # check on the rank's first owned block of tile columns (local data only)
    d = torch.empty_like(A0)
    rk.download(d.data_ptr())
    j0 = own[0]
    j1 = j0
    while j1 + 1 in own and j1 + 1 - j0 < rk.block:
        j1 += 1
    gram = gram_check(torch, d, A0, j0 * 32, (j1 + 1) * 32)
    t = torch.tensor([gram], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gram = t.item()
    del d

    # e2e: owned columns travel host -> device -> host around the factorisation
    cols = torch.tensor([c for j in own for c in range(j * 32, (j + 1) * 32)], device="cuda")
    h_in = A0[cols].cpu().pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    dwork = torch.empty((len(cols), n), device="cuda", dtype=torch.float32)
    mat = None
    e2e_s = 0.0
    e2e_steps = max(1, min(args.steps, 2))
    full = torch.empty((n, n), device="cuda", dtype=torch.float32)
    for it in range(e2e_steps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        dwork.copy_(h_in, non_blocking=True)                     # H2D of the owned columns
        full[cols] = dwork
        rk.upload(full.data_ptr())
        rk.reset(stream)
        torch.cuda.synchronize()
        dist.barrier()
        rk.factor(stream)
        rk.download(full.data_ptr())
        h_out.copy_(full[cols], non_blocking=True)               # D2H of the owned columns
        torch.cuda.synchronize()
        dist.barrier()
        if it > 0:
            e2e_s += time.perf_counter() - t0
    t = torch.tensor([e2e_s / e2e_steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = flops / t.item() * 1e-9
    del full, dwork

    # the same workload on ONE GPU (rank 0, one step) so the strong-scaling ratio can be read off this line
    single = None
    if rank == 0 and not args.no_single_ref:
        plan = q.Plan(n, n)
        d = A0.clone()
        plan.factor(d.data_ptr(), n, stream)
        torch.cuda.synchronize()
        d.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.factor(d.data_ptr(), n, stream); e1.record()
        torch.cuda.synchronize()
        single = flops / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        plan.close()
        del d
    dist.barrier()
    if rank != 0:
        rk.close()
        return
    peaks, peak_src = measured_peaks()
    fp32 = C.c_double(0)
    q.check(L.b200qr_measure_fp32_peak(C.byref(fp32), 20000))
    achieved_tf = value * 1e-3
    variant = rk.variant
    wpeaks = {k2: (v * world if isinstance(v, (int, float)) else v) for k2, v in peaks.items()}
    roofline = roofline_of(variant, achieved_tf, fp32.value * 1e-3 * world, wpeaks, peak_src + " x %d GPUs" % world, n, ms_step)
    roofline["kernel"] += " -- one launch per rank per step"
    how = ("operand tiles (T factor, reflector block and its transpose, split for 3xTF32) broadcast by the panel owner's TMA "
           "stores into every rank's workspace over NVLink, from inside the persistent kernel" if variant == 5 else
           "reflector tiles read from the owner over NVLink inside the persistent kernel")
    line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform[-1,1), torch cuda generator seed 5",
            "config": {"workload": "%dx%d square tiled QR, tile size 32, flat-tree TS kernels" % (n, n),
                       "parallelism": "1-D block-cyclic tile columns (blocks of %d tile columns) over %d GPUs; %s"
                                      % (rk.block, world, how), "kernel_variant": variant,
                       "l2": "input restored from a pristine copy before every timed step",
                       "persistent_grid_ctas_per_rank": rk.grid,
                       "single_gpu_same_workload_gflops": single},
            "roofline": roofline, "cpu_baseline": None,
            "e2e": {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n, "d2h_bytes_per_step": 4 * n * n,
                    "api": "b200qr_dist_* with pinned host copies of the owned columns", "steps": e2e_steps},
            "gpu_launches": args.steps * world, "clocks": clocks, "check": {"gram_rel_err_owned_block_max": gram}}
    print(json.dumps(line), flush=True)
    rk.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--north-size", type=int, default=32768,
                    help="N = 1: also factor this size once (BASELINE.json's 32768 target) and report it; 0 = skip")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="CPU baseline sample size (square)")
    ap.add_argument("--ref-size", type=int, default=0)
    ap.add_argument("--multi-size", type=int, default=32768, help="matrix size of the N > 1 (block-cyclic) workload")
    ap.add_argument("--no-single-ref", action="store_true", help="N > 1: skip the same-size single-GPU reference step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
