#!/usr/bin/env python
"""bench.py -- QR GFLOP/s (flops = 2mn^2 - 2n^3/3) of the B200 tiled QR, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size n] [--extras list|none]

A step = one tiled QR factorisation of a synthetic 32768 x 32768 float32 matrix (BASELINE.json configs[4], the
configuration the north-star target is quoted on; it fits one GPU) with the reference's tile size 32 -- the SAME
workload at N = 1, 2, 4, 8, so the driver's scaling efficiency compares like with like (strong scaling).  `value` is
measured with the matrix resident in HBM (CUDA events on the launching stream around the factorisation only, like the
reference's own "GPU: x ms" line, gpucalc.cu:1635-1647); `e2e` is the same metric through the host-facing C ABI
(cudaQRTask on a pinned HOST buffer at N = 1, b200qr_dist_* with pinned host copies of the owned tile columns at N > 1),
H2D and D2H copies inside the timed region.  Between timed steps the 4 GiB input is restored from a pristine copy,
which also evicts the 126 MB L2.

The N = 1 line also carries the other BASELINE.json configurations as extra keys under "configs", each with its own
roofline object and accuracy check: the 8192^2 tile-size sweep (configs[1]), 10 000 batched 256^2 problems (configs[2])
and the 4 194 304 x 256 TSQR (configs[3]); the N > 1 line carries the TSQR with its rows split over the N ranks.
`--impl reference` times the reference's own CPU path (oracle/_ref, 8 pthreads) on a bounded sample.  One JSON line
on stdout (rank 0).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "QR GFLOP/s (2mn^2-2n^3/3)"
EPS32 = 2.0 ** -23


def qr_flops(m, n):
    return 2.0 * m * n * n - 2.0 * n ** 3 / 3.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, shape):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed ncu capture of that kernel and
    shape (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of the run named there); None when
    that shape has not been captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get("%s@%s" % (kernel, shape))
        return (e["dram_bytes_read"] + e["dram_bytes_write"], e.get("source")) if e else (None, None)
    except Exception:
        return None, None


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


# ------------------------------------------------------------------------------------------------ CPU reference arm
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


def cpu_baseline_obj(sample_n):
    sec, kind, cores = cpu_reference_run(sample_n, reps=1)
    return {"value": qr_flops(sample_n, sample_n) / sec * 1e-9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": "%dx%d RANDZO matrix (the reference's own generator), b=32, one run of taskQRP_threads (%.1f s), "
                      "-O3 -march=x86-64-v3, 8 pthreads = the reference's NUMTHREADS (qrdecomp.c:21); host has %d cpus. "
                      "The 32768^2 workload itself is out of the reference's reach: its scheduler scans the whole p x q task "
                      "grid under one mutex per fetch (gridscheduler.c:259-300)" % (sample_n, sample_n, sec, os.cpu_count())}


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
            "config": {"workload": "%dx%d square tiled QR, tile size 32 (reference default), flat-tree TS kernels"
                                   % (args.size, args.size),
                       "sample": "each step = one %dx%d factorisation by the reference's CPU path: a bounded sample of the "
                                 "workload (the reference's O(p*q)-scan scheduler needs hours at 32768 and has no multi-GPU "
                                 "path); GFLOP/s is size-normalised, and the reference gets SLOWER per flop as the task grid "
                                 "grows (SURVEY section 6)" % (n, n)},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ checks / rooflines
def gram_check(torch, d, A0, c0, c1):
    """|R^T R - A^T A| / |A^T A| restricted to matrix columns [c0, c1): needs only those columns of
    the factored matrix d (R on/above the diagonal) and of the input A0.  Storage is (n, ld)."""
    rows = torch.arange(c1, device=d.device)[:, None]
    cols = torch.arange(c0, c1, device=d.device)[None, :]
    R = torch.where(rows <= cols, d[c0:c1, :c1].t().double(), torch.zeros((), dtype=torch.float64, device=d.device))
    Ad = A0[c0:c1].t().double()
    G2 = Ad.t() @ Ad
    return (torch.linalg.norm(R.t() @ R - G2) / torch.linalg.norm(G2)).item()


def gram_tall(torch, R, A0, m, n):
    """|R^T R - A^T A| / |A^T A| for a tall matrix stored (n, m); R stored (n, n) column-major."""
    Rm = R.t().double()
    G = torch.zeros((n, n), dtype=torch.float64, device=A0.device)
    for c in range(0, m, 1 << 20):
        blk = A0[:, c:c + (1 << 20)].double()
        G += blk @ blk.t()
    return (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item()


def fp32_peak_tf(L, q):
    fp32 = C.c_double(0)
    q.check(L.b200qr_measure_fp32_peak(C.byref(fp32), 20000))
    return fp32.value * 1e-3


def roofline_square(variant, achieved_tf, fp32_tf, peaks, peak_src, n, ms_step, ngpu=1):
    """The factorisation is one persistent kernel per GPU, so the kernel's roofline is the step's.  Variants 1 / 2 apply
    the reflectors on the FP32 SIMT pipe (bound: the FFMA rate, measured in this process).  Variant 5 runs the trailing
    update as 3xTF32 tcgen05 products: the bound is the tensor pipe at a sixth of the measured dense bf16 rate (TF32 at
    half the bf16 rate, three passes per product); the fraction of the FP32 SIMT peak -- the north-star's '% of B200
    peak for the working precision' -- is reported next to it."""
    kern = "tiledqr_persistent_tc3" if variant == 5 else "tiledqr_persistent"
    traffic, tsrc = ncu_traffic(kern, "%dx%d" % (n, n)) if ngpu == 1 else (None, None)
    common = {"achieved": achieved_tf, "unit": "TFLOP/s", "traffic": traffic, "traffic_source": tsrc,
              "algorithmic_bytes": 8.0 * n * n, "algorithmic_flops": qr_flops(n, n),
              "frac_of_fp32_simt_peak": achieved_tf / (fp32_tf * ngpu), "fp32_simt_peak": fp32_tf * ngpu,
              "frac_of_bf16_tensor_peak": achieved_tf / (peaks["bf16_tflops"] * ngpu),
              "hbm_algorithmic_gbs": 8.0 * n * n / (ms_step * 1e-3) * 1e-9}
    if variant == 5:
        peak = peaks["bf16_tflops"] / 6.0 * ngpu
        common.update({"bound": "tensor", "peak": peak, "frac": achieved_tf / peak,
                       "kernel": "tiledqr_persistent_tc3 (the only kernel of a step%s)" % (", one launch per rank" if ngpu > 1 else ""),
                       "peak_source": "%s bf16_tflops / 6 (3xTF32: TF32 at half the bf16 rate, three passes)%s; FP32 SIMT "
                                      "peak from the FFMA microbench run in this process"
                                      % (peak_src, " x %d GPUs" % ngpu if ngpu > 1 else "")})
    else:
        common.update({"bound": "fp32-simt", "peak": fp32_tf * ngpu, "frac": achieved_tf / (fp32_tf * ngpu),
                       "kernel": "tiledqr_persistent (the only kernel of a step)",
                       "peak_source": "FFMA microbench run in this process (b200qr_measure_fp32_peak); %s carries "
                                      "only HBM and bf16-tensor peaks" % peak_src})
    return common


def roofline_hbm(kernel, shape, bytes_alg, flops_alg, ms, fp32_tf, peaks, peak_src, ngpu=1):
    """Panel-kernel-bound configurations (batched, TSQR): reported against the HBM roofline on their algorithmic bytes
    (read A once, write R + V once: 8 m n), as the north-star asks for panel kernels, with the FP32 fraction next to it
    (both sit near the FP32 ridge, the smaller of the two bounds is the one that binds)."""
    gbs = bytes_alg / (ms * 1e-3) * 1e-9
    tf = flops_alg / (ms * 1e-3) * 1e-12
    traffic, tsrc = ncu_traffic(kernel, shape) if ngpu == 1 else (None, None)
    t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9 * ngpu)
    t_fp32 = flops_alg / (fp32_tf * 1e12 * ngpu)
    return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"] * ngpu, "unit": "GB/s", "frac": gbs / (peaks["hbm_gbs"] * ngpu),
            "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops_alg,
            "kernel": kernel, "peak_source": peak_src + (" x %d GPUs" % ngpu if ngpu > 1 else ""),
            "achieved_tflops": tf, "frac_of_fp32_simt_peak": tf / (fp32_tf * ngpu),
            "binding_term": "fp32-simt" if t_fp32 > t_hbm else "hbm",
            "min_ms_hbm": t_hbm * 1e3, "min_ms_fp32_simt": t_fp32 * 1e3}


def timed(torch, fn, restore, steps, warmup):
    for _ in range(warmup):
        restore(); fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        restore()
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps


# ------------------------------------------------------------------------------------------------ extra configurations
def extra_sweep(torch, q, L, fp32_tf, peaks, peak_src, steps, n=8192):
    """BASELINE.json configs[1]: 8192 x 8192 on one GPU, tile size 32 / 64 / 128.  32 is the reference-compatible size
    (drop-in output layout); 64 and 128 are the documented non-default mode with their own reflector layout."""
    from qr_checks import probe_accuracy
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    d = torch.empty_like(A0)
    rows = []
    flops = qr_flops(n, n)
    for b in (32, 64, 128):
        row = {"tile": b}
        try:
            if b == 32:
                plan = q.Plan(n, n)
                fn = lambda: plan.factor(d.data_ptr(), n, stream)
                ms = timed(torch, fn, lambda: d.copy_(A0), steps, 3)
                back, orth = probe_accuracy(torch, plan, d, A0, n, n, 32)
                row.update({"kernel_variant": plan.variant, "backward_err_n_eps": back, "orthogonality_n_eps": orth})
                tf = flops / ms * 1e-9
                row["roofline"] = roofline_square(plan.variant, tf, fp32_tf, peaks, peak_src, n, ms)
                plan.close()
            else:
                if not hasattr(q, "TilePlan"):
                    raise RuntimeError("tile size %d not built" % b)
                plan = q.TilePlan(n, n, b)
                fn = lambda: plan.factor(d.data_ptr(), n, stream)
                ms = timed(torch, fn, lambda: d.copy_(A0), max(1, min(steps, 3)), 1)
                back, orth = probe_accuracy(torch, plan, d, A0, n, n, 128)
                row.update({"backward_err_n_eps": back, "orthogonality_n_eps": orth})
                tf = flops / ms * 1e-9
                row["roofline"] = {"bound": "fp32-simt", "achieved": tf, "peak": fp32_tf, "unit": "TFLOP/s", "frac": tf / fp32_tf,
                                   "traffic": None, "kernel": "tileqr_step_kernels<%d> (host-ordered, one panel + one update launch per step)" % b}
                plan.close()
            row.update({"ms": ms, "value": flops / ms * 1e-6, "unit": "GFLOP/s", "gram_rel_err": gram_check(torch, d, A0, 0, 1024)})
        except Exception as e:
            row["error"] = str(e)[:200]
        rows.append(row)
    return {"workload": "%dx%d square tiled QR on 1 GPU, tile-size sweep" % (n, n), "rows": rows}


def extra_batched(torch, q, L, fp32_tf, peaks, peak_src, steps, nprob=10000, m=256, n=256):
    """BASELINE.json configs[2]: 10 000 independent 256 x 256 problems, each factored exactly like cudaQRTask would."""
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((nprob, n, m), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    fl = C.c_longlong(0)
    q.check(L.b200qr_batched_workspace_floats(m, n, nprob, C.byref(fl)))
    tau = torch.zeros(fl.value, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    fn = lambda: q.check(L.b200qr_batched_factor(C.c_void_p(d.data_ptr()), m, n, m, m * n, nprob, C.c_void_p(tau.data_ptr()),
                                                 C.c_void_p(ctr.data_ptr()), C.c_void_p(stream)))
    ms = timed(torch, fn, lambda: d.copy_(A0), steps, 3)
    flops = nprob * qr_flops(m, n)
    # checks: R^T R = A^T A for 16 problems spread over the batch; problem 0 bit-equal to the single-matrix entry point
    errs = []
    for b in range(0, nprob, max(1, nprob // 16)):
        R = torch.triu(d[b].t().double()); Ad = A0[b].t().double()
        errs.append((torch.linalg.norm(R.t() @ R - Ad.t() @ Ad) / torch.linalg.norm(Ad.t() @ Ad)).item())
    plan = q.Plan(m, n)
    one = A0[0].clone()
    plan.factor(one.data_ptr(), m, stream)
    torch.cuda.synchronize()
    same = bool(torch.equal(one, d[0]))
    plan.close()
    return {"workload": "batched %d x (%dx%d), tile size 32, reference layout per problem, 1 GPU" % (nprob, m, n),
            "ms": ms, "value": flops / ms * 1e-6, "unit": "GFLOP/s", "gram_rel_err_max_of_16": max(errs),
            "problem0_equals_single_matrix_path": same,
            "roofline": roofline_hbm("batched_warp_kernel", "%dx%dx%d" % (nprob, m, n), 8.0 * nprob * m * n, flops, ms,
                                     fp32_tf, peaks, peak_src)}


def extra_tsqr(torch, q, L, fp32_tf, peaks, peak_src, steps, m=4194304, n=256):
    """BASELINE.json configs[3] on one GPU."""
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.rand((n, m), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    t = q.Tsqr(m, n)
    R = torch.zeros((n, n), device="cuda")
    fn = lambda: t.factor(d.data_ptr(), m, R.data_ptr(), stream)
    ms = timed(torch, fn, lambda: d.copy_(A0), steps, 3)
    flops = qr_flops(m, n)
    err = gram_tall(torch, R, A0, m, n)
    out = {"workload": "TSQR %dx%d on 1 GPU (%s)" % (m, n, t.describe() if hasattr(t, "describe") else "streaming engine"),
           "ms": ms, "value": flops / ms * 1e-6, "unit": "GFLOP/s", "gram_rel_err": err,
           "r_lower_triangle_zero": bool((torch.tril(R.t(), -1) == 0).all().item()),
           "roofline": roofline_hbm(t.kernel_name if hasattr(t, "kernel_name") else "tsqr_strip_kernel", "%dx%d" % (m, n),
                                    8.0 * m * n, flops, ms, fp32_tf, peaks, peak_src)}
    t.close()
    return out


# ------------------------------------------------------------------------------------------------ N = 1
def run_b200_single(args):
    import torch
    import gpu_tiled_qr_decomposition_b200 as q
    from qr_checks import probe_accuracy
    q.build()
    L = q.lib()
    n = args.size
    flops = qr_flops(n, n)
    stream = torch.cuda.current_stream().cuda_stream
    peaks, peak_src = measured_peaks()
    fp32_tf = fp32_peak_tf(L, q)
    extras = [] if args.extras == "none" else args.extras.split(",")

    line = {}
    if "headline" in args.only:
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
        back, orth = probe_accuracy(torch, plan, d, A0, n, n, 32)
        variant, grid, tile_tasks, macro = plan.variant, plan.grid, plan.tile_tasks, plan.macro_tasks

        # ---- e2e: the legacy host entry point on a pinned host buffer
        os.environ["QRB200_QUIET"] = "1"
        h = torch.empty((n, n), dtype=torch.float32).pin_memory()
        h0 = A0.cpu()
        dref = d.cpu()
        plan.close()
        del d, A0
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.steps, 2))
        h.copy_(h0)
        L.cudaQRTask(C.c_void_p(h.data_ptr()), n, n, n, 128)          # warm-up (plan, device buffer, pinned pages)
        e2e_s = 0.0
        for _ in range(e2e_steps):
            h.copy_(h0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            L.cudaQRTask(C.c_void_p(h.data_ptr()), n, n, n, 128)
            e2e_s += time.perf_counter() - t0
        e2e_val = flops / (e2e_s / e2e_steps) * 1e-9
        e2e_err = float((h - dref).abs().max())                        # whole matrix, vs the device-resident path
        L.b200qr_release_cache()
        del h, h0, dref

        achieved_tf = flops / (ms_step * 1e-3) * 1e-12
        line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform[-1,1), torch cuda generator seed 5",
                "config": {"workload": "%dx%d square tiled QR, tile size 32 (reference default), flat-tree TS kernels"
                                       % (n, n), "parallelism": "1 GPU",
                           "l2": "input restored from a pristine %d MiB copy before every timed step (> 126 MB L2)"
                                 % (4 * n * n >> 20),
                           "persistent_grid_ctas": grid, "kernel_variant": variant, "tile_tasks": tile_tasks,
                           "macro_tasks": macro},
                "roofline": roofline_square(variant, achieved_tf, fp32_tf, peaks, peak_src, n, ms_step),
                "e2e": {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n,
                        "d2h_bytes_per_step": 4 * n * n, "api": "cudaQRTask(host float*, m, n, ldm, maxblocks) on pinned "
                        "host memory (workspace cached between calls of one shape; copies overlapped with the kernel: tile "
                        "columns stream in behind progress words, finished columns stream out)", "steps": e2e_steps,
                        "max_abs_diff_vs_device_path": e2e_err},
                "gpu_launches": args.steps * 1, "clocks": clocks,
                "check": {"gram_rel_err_first_2048_cols": gram, "backward_err_n_eps": back, "orthogonality_n_eps": orth,
                          "bound_n_eps": 10.0, "how": "fp64 products with A and R, Q applied by b200qr_apply_q on 32 random probes"}}
    cfgs = {}
    esteps = max(1, min(args.steps, 5))
    for name, fn in (("tile_sweep_8192", extra_sweep), ("batched_10000x256", extra_batched), ("tsqr_4194304x256", extra_tsqr)):
        key = name.split("_")[0] if name != "tile_sweep_8192" else "sweep"
        if key in extras or name in extras:
            try:
                cfgs[name] = fn(torch, q, L, fp32_tf, peaks, peak_src, esteps)
            except Exception as e:                                    # e.g. not enough free memory on a shared box
                cfgs[name] = {"error": str(e)[:300]}
            torch.cuda.empty_cache()
    if cfgs:
        line["configs"] = cfgs
    if "headline" in args.only and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_obj(args.cpu_sample)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ N > 1
def tsqr_multi(torch, dist, q, rank, world, steps, fp32_tf, peaks, peak_src, m=4194304, n=256):
    """BASELINE.json configs[3] with the rows split over the ranks: local reduction, all-gather of the W R factors
    (the path's one exchange step), redundant top reduction on every rank."""
    m_local = m // world
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    A0 = torch.rand((n, m_local), generator=g, device="cuda") * 2 - 1
    d = torch.empty_like(A0)
    R = torch.zeros((n, n), device="cuda")
    td = q.TsqrDistC(m_local, n, rank, world)                 # the C ABI: exchange over peer memory inside the stream
    blobs = [None] * world
    dist.all_gather_object(blobs, td.export())
    td.connect(blobs)
    ms_tot = 0.0
    for it in range(steps + 2):
        d.copy_(A0)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); td.factor(d.data_ptr(), m_local, R.data_ptr(), stream); e1.record()
        torch.cuda.synchronize(); dist.barrier()
        if it >= 2:
            ms_tot += e0.elapsed_time(e1)
    t = torch.tensor([ms_tot / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # checks: same R everywhere; R^T R = sum over ranks of A_r^T A_r
    Rs = [torch.empty_like(R) for _ in range(world)]
    dist.all_gather(Rs, R)
    same = all(torch.equal(Rs[0], x) for x in Rs)
    G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for c in range(0, m_local, 1 << 20):
        blk = A0[:, c:c + (1 << 20)].double()
        G += blk @ blk.t()
    dist.all_reduce(G)
    Rm = R.t().double()
    err = (torch.linalg.norm(Rm.t() @ Rm - G) / torch.linalg.norm(G)).item()
    td.close()
    flops = qr_flops(m, n)
    return {"workload": "TSQR %dx%d, rows split over %d GPUs (%d rows each; local engine: %s), R factors exchanged over peer "
                        "memory inside the stream (b200qr_tsqr_dist_*: no collective library on the data path) + redundant top "
                        "reduction" % (m, n, world, m_local, td.mode), "ms": ms, "value": flops / ms * 1e-6, "unit": "GFLOP/s",
            "same_R_on_all_ranks": bool(same), "gram_rel_err": err, "exchange_bytes_per_rank": 4 * n * n * (world - 1),
            "roofline": roofline_hbm("tsqr local reduction + top reduction", "%dx%d" % (m, n), 8.0 * m * n, flops, ms, fp32_tf,
                                     peaks, peak_src, world)}


def run_b200_multi(args, rank, world, local_rank):
    """N > 1: one 32768 x 32768 matrix (BASELINE.json configs[4]), tile columns dealt to the ranks
    1-D block-cyclically; every rank runs the persistent kernel on its own GPU; the panel broadcast happens
    inside it (TMA stores into the peers' workspaces over NVLink).  Strong scaling: total work is fixed."""
    import torch
    import torch.distributed as dist
    import gpu_tiled_qr_decomposition_b200 as q
    q.build()
    L = q.lib()
    n = args.size
    flops = qr_flops(n, n)
    stream = torch.cuda.current_stream().cuda_stream
    m = n

    rk = q.DistRank(n, n, rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, rk.export())
    rk.connect(blobs)
    g = torch.Generator(device="cuda").manual_seed(5)            # same matrix on every rank
    A0 = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    own = rk.owned_tile_cols()
    # owned tile columns as contiguous runs [j0, j1): a run is one contiguous range of the column-major matrix
    runs = []
    for j in own:
        if runs and runs[-1][1] == j:
            runs[-1][1] = j + 1
        else:
            runs.append([j, j + 1])

    def one_step():
        rk.upload(A0.data_ptr())                                 # restore (device-to-device, evicts L2)
        rk.reset(stream)                                         # synchronous: the words ARE zero when it returns
        dist.barrier()                                           # every rank has reset its progress words
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rk.factor(stream)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    ms = sum(one_step() for _ in range(args.steps))
    dist.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = flops / (ms_step * 1e-3) * 1e-9

    # ---- checks.  (1) Gram identity on every owned run's first block, max over ranks.  (2) The whole factored matrix,
    # assembled from the owners, against the single-GPU factorisation of the same engine on rank 0: every element.
    d = torch.empty_like(A0)
    rk.download(d.data_ptr())
    gram = gram_check(torch, d, A0, runs[0][0] * 32, min(runs[0][1], runs[0][0] + 4) * 32)
    t = torch.tensor([gram], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gram = t.item()
    mask = torch.zeros(n // 32, device="cuda", dtype=torch.bool)
    mask[torch.tensor(own, device="cuda")] = True
    d.view(n // 32, 32 * m).mul_(mask[:, None].float())          # zero the tile columns this rank does not own
    dist.reduce(d, dst=0)                                        # rank 0: the assembled result (NCCL sum of disjoint parts)
    single, ident, back, orth = None, None, None, None
    if rank == 0 and not args.no_single_ref:
        from qr_checks import probe_accuracy
        plan = q.Plan(n, n)
        s1 = A0.clone()
        plan.factor(s1.data_ptr(), n, stream)
        torch.cuda.synchronize()
        ident = {"single_gpu_variant": plan.variant, "identical": bool(torch.equal(s1, d)),
                 "max_abs_diff": float((s1 - d).abs().max().item())}
        s1.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.factor(s1.data_ptr(), n, stream); e1.record()
        torch.cuda.synchronize()
        single = flops / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        back, orth = probe_accuracy(torch, plan, s1, A0, n, n, 32)
        plan.close()
        del s1
    del d
    torch.cuda.empty_cache()
    dist.barrier()

    # ---- e2e: b200qr_dist_reset_host / b200qr_dist_factor_host on pinned host memory -- the rank's kernel starts at once,
    # its owned tile columns stream in behind it (progress words gate the step-0 tasks) and every finished block of
    # columns streams back while the kernel works on the rest.  The host buffer holds the rank's columns only (packed).
    h_in = torch.cat([A0[j0 * 32:j1 * 32] for j0, j1 in runs]).cpu().pin_memory()
    h_work = torch.empty_like(h_in).pin_memory()
    e2e_steps = max(1, min(args.steps, 2))
    e2e_s = 0.0
    e2e_same = True
    for it in range(e2e_steps + 1):
        h_work.copy_(h_in)                                           # (host to host: restores the input, not timed)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        rk.reset_host()
        dist.barrier()                                               # every rank's words say "input not here yet"
        rk.factor_host(h_work.data_ptr(), m, packed=True)
        dist.barrier()
        if it > 0:
            e2e_s += time.perf_counter() - t0
    # the columns that came back are the ones the device-resident run produced
    dchk = torch.empty((n, m), device="cuda", dtype=torch.float32)
    rk.download(dchk.data_ptr())
    e2e_same = bool(torch.equal(torch.cat([dchk[j0 * 32:j1 * 32] for j0, j1 in runs]).cpu(), h_work))
    del dchk
    same_t = torch.tensor([1 if e2e_same else 0], device="cuda")
    dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
    e2e_same = bool(same_t.item())
    t = torch.tensor([e2e_s / e2e_steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = flops / t.item() * 1e-9
    variant, block, grid = rk.variant, rk.block, rk.grid
    rk.close()
    del A0, h_in, h_work
    torch.cuda.empty_cache()

    peaks, peak_src = measured_peaks()
    fp32_tf = fp32_peak_tf(L, q)
    tsq = None
    if "tsqr" in (args.extras.split(",") if args.extras != "none" else []):
        try:
            tsq = tsqr_multi(torch, dist, q, rank, world, max(1, min(args.steps, 5)), fp32_tf, peaks, peak_src)
        except Exception as e:
            tsq = {"error": str(e)[:300]}
    dist.barrier()
    if rank != 0:
        return
    achieved_tf = value * 1e-3
    roofline = roofline_square(variant, achieved_tf, fp32_tf, peaks, peak_src, n, ms_step, world)
    how = ("operand tiles (T factor, reflector block and its transpose, split for 3xTF32) broadcast by the panel owner's TMA "
           "stores into every rank's workspace over NVLink, from inside the persistent kernel" if variant == 5 else
           "reflector tiles read from the owner over NVLink inside the persistent kernel")
    line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform[-1,1), torch cuda generator seed 5",
            "config": {"workload": "%dx%d square tiled QR, tile size 32 (reference default), flat-tree TS kernels" % (n, n),
                       "parallelism": "1-D block-cyclic tile columns (blocks of %d tile columns) over %d GPUs; %s"
                                      % (block, world, how), "kernel_variant": variant,
                       "l2": "input restored from a pristine copy before every timed step",
                       "persistent_grid_ctas_per_rank": grid,
                       "single_gpu_same_workload_gflops": single},
            "roofline": roofline,
            "cpu_baseline": None if args.no_cpu else cpu_baseline_obj(args.cpu_sample),
            "e2e": {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n, "d2h_bytes_per_step": 4 * n * n,
                    "api": "b200qr_dist_reset_host / b200qr_dist_factor_host (C ABI) on pinned host memory, one call per rank: "
                           "the owned tile columns stream in behind the running kernel and finished blocks stream back "
                           "while it works on the rest", "steps": e2e_steps, "equals_device_path": e2e_same},
            "gpu_launches": args.steps * world, "clocks": clocks,
            "check": {"gram_rel_err_owned_block_max": gram, "whole_matrix_vs_single_gpu": ident,
                      "backward_err_n_eps": back, "orthogonality_n_eps": orth, "bound_n_eps": 10.0}}
    if tsq is not None:
        line["configs"] = {"tsqr_4194304x256": tsq}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=32768, help="matrix size of the headline workload (square)")
    ap.add_argument("--extras", default="sweep,batched,tsqr",
                    help="other BASELINE configurations reported under 'configs' (comma list of sweep,batched,tsqr; or none)")
    ap.add_argument("--only", default="headline", help="'headline' (default: headline + extras) or 'extras' (profiling runs)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="CPU baseline sample size (square)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--ref-size", type=int, default=0)
    ap.add_argument("--no-single-ref", action="store_true", help="N > 1: skip the same-size single-GPU reference step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        try:
            run_b200_multi(args, rank, world, local_rank)
        finally:
            dist.destroy_process_group()
    else:
        run_b200_single(args)


if __name__ == "__main__":
    main()
