/*
 * qrb200_probe.h -- diagnostics library libqrb200_probe.so (NOT part of the product ABI, include/qrb200.h).
 * Micro-measurements that shaped the kernels: panel-kernel latency per tile, tcgen05 building blocks and issue rates.
 * Built next to libqrb200.so by the package Makefile from csrc/probes; loaded by tools/probe_*.py and tests/test_gpu_umma.py.
 */
#ifndef QRB200_PROBE_H
#define QRB200_PROBE_H
#ifdef __cplusplus
extern "C" {
#endif
/* SM cycles per 32x32 tile of the one-warp panel kernels (geqrt_warp / tsqrt_warp_v, csrc/qr_device.cuh): nwarps (<= 16)
 * independent chains share one SM; busy != 0 adds eight FFMA-bound warps (then nwarps <= 8).  cycles[2w] = GEQRT,
 * cycles[2w + 1] = TSQRT of warp w; out3 (optional, 3 x 1024 floats) = warp 0's GEQRT tile, the triangle after its last
 * TSQRT and that TSQRT's V2, column-major (input: a fixed pseudo-random tile).  0 = ok. */
int b200qr_probe_panel(int nwarps, int busy, int reps, long long *cycles, float *out3);
/* SM cycles per tile row (16 tiles x 32 reflectors) of the SIMT update inner loop on one SM, nc = 1 / 2 / 4 matrix columns
 * per thread (16 / nc warps), operands in shared memory: the rate the sweep consumers can reach at best.  0 = ok. */
int b200qr_probe_sweep(int nc, int iters, long long *cycles);
/* One read + write pass over an n x n float32 matrix in the sweep order of the tensor-core engine (warp = tile, lane = one
 * 128-byte line): packed = 0 column-major (lines ld*4 bytes apart), 1 = every 32 x 32 tile contiguous.  ms_out = best of 3. */
int b200qr_probe_dram_pattern(int n, int packed, int ctas_per_sm, float *ms_out);
/* tcgen05 building-block check (csrc/probes/umma_probe.cu): hA 32 x 128 column-major, hV 32 x 32 column-major,
 * hW 128 x 32 row-major or NULL; hZ <- A^T V (128 x 32 row-major, 3xTF32 on the tensor cores);
 * hD2 <- A^T - W V^T when hW is given. */
int b200qr_probe_umma(const float *hA, const float *hV, const float *hW, float *hZ, float *hD2);
int b200qr_probe_umma_mode(const float *hA, const float *hV, const float *hW, float *hZ, float *hD2, int mode);
int b200qr_probe_umma_rate(long long *out8, int reps);   /* cycles of reps*4 tcgen05.mma (tf32, M=128, K=8): N=32,64,128,256 x {1,2} accumulators */
int b200qr_probe_umma_rate2(long long *out2, int reps);  /* one vs two issuing threads, reps*4 MMAs each (N = 32) */
#ifdef __cplusplus
}
#endif
#endif
