// microbench.cu -- roofline denominators that MEASURED_PEAKS.json does not carry.
//
// The tiled QR computes in float32 on the SIMT FFMA pipe; the driver-measured peaks are an
// HBM copy and a bf16 tensor GEMM.  b200qr_measure_fp32_peak() times a register-resident FFMA
// loop on every SM so bench.py can report the factorisation against the FP32 rate the chip
// actually sustains at the clocks of the run (SURVEY 8d asks for exactly this measurement).
#include "../../include/qrb200.h"
#include <cuda_runtime.h>

namespace qrb200 {

// 16 independent accumulator chains per thread, operands in registers; 2 flops per FFMA.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *sink, int iters, float a, float b)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678f) sink[0] = s;          // keeps the loop alive, never true in practice
}

}  // namespace qrb200

extern "C" int b200qr_measure_fp32_peak(double *gflops, int iters)
{
    if (!gflops) return B200QR_EINVAL;
    if (iters <= 0) iters = 20000;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return B200QR_ECUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return B200QR_ECUDA;
    float *sink = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) return B200QR_ECUDA;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {          // first rep warms up clocks
        cudaEventRecord(e0, 0);
        qrb200::ffma_peak_kernel<<<blocks, threads>>>(sink, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return B200QR_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    *gflops = flops / (best * 1e-3) * 1e-9;
    return B200QR_OK;
}

// ---------------------------------------------------------------------------------------------
// Panel-kernel latency probe: one warp runs GEQRT / TSQRT `reps` times on a tile kept in shared
// memory and reports SM cycles per tile.  mode 0: the warp is alone on its SM; mode 1: eight
// more warps of the CTA run an FFMA loop (what the sweep consumers do to the panel warp).
#include "qr_device.cuh"
namespace qrb200 {
__global__ void __launch_bounds__(288) panel_probe_kernel(const float *tile_in, float *tile_out, long long *cycles,
                                                          int reps, int mode, volatile int *stop)
{
    __shared__ __align__(16) float Rs[TB * 33];
    __shared__ __align__(16) float xs[2 * TB];
    __shared__ __align__(16) float ts[TB * TB];
    __shared__ __align__(16) float ts2[TB * TB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        float a[TB], b[TB];
        load_col_gmem(a, tile_in + lane * TB);
        store_col_smem(ts + lane * TB, a);
        __syncwarp();
        long long t0 = clock64();
        float acc = 0.f;
        for (int r = 0; r < reps; ++r) {
            load_col_smem(a, ts + lane * TB);
            acc += geqrt_warp(a, xs, lane);
        }
        long long t1 = clock64();
#pragma unroll
        for (int r = 0; r < TB; ++r) Rs[r * 33 + lane] = a[r];
        __syncwarp();
        for (int r = 0; r < reps; ++r) {
            load_col_smem(b, ts + lane * TB);
            acc += tsqrt_warp_v<false>(b, Rs, xs, ts2, lane);
        }
        long long t2 = clock64();
        store_col_gmem(tile_out + lane * TB, b);
        if (lane == 0) { cycles[0] = (t1 - t0) / reps; cycles[1] = (t2 - t1) / reps; tile_out[0] += acc; *stop = 1; }
    } else if (mode == 1) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = (float)(threadIdx.x + i);
        while (*stop == 0) {
#pragma unroll
            for (int u = 0; u < 64; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], 1.0000001f, 1e-7f);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc[i];
        if (s == 12345.678f) tile_out[1] = s;
    }
}
}  // namespace qrb200

extern "C" int b200qr_probe_panel(long long *cycles4, int reps)
{
    if (!cycles4 || reps <= 0) return B200QR_EINVAL;
    float h[32 * 32];
    for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 37 % 201) - 100) / 100.0f;
    float *din = nullptr, *dout = nullptr;
    long long *dc = nullptr;
    int *dstop = nullptr;
    if (cudaMalloc(&din, 4096) || cudaMalloc(&dout, 4096) || cudaMalloc(&dc, 16) || cudaMalloc(&dstop, 4)) return B200QR_ECUDA;
    cudaMemcpy(din, h, 4096, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dstop, 0, 4);
        qrb200::panel_probe_kernel<<<1, 288>>>(din, dout, dc, reps, mode, dstop);
        if (cudaDeviceSynchronize() != cudaSuccess) return B200QR_ECUDA;
        cudaMemcpy(cycles4 + 2 * mode, dc, 16, cudaMemcpyDeviceToHost);
    }
    cudaFree(din); cudaFree(dout); cudaFree(dc); cudaFree(dstop);
    return B200QR_OK;
}

// ---------------------------------------------------------------------------------------------
// How many independent one-warp TSQRT chains does an SM run at full speed?  `nwarps` warps of one CTA (<= 16) each run
// tsqrt_warp_v `reps` times on their own tile in shared memory; out[w] = SM cycles per tile seen by warp w.  (The
// round-2 plan for tall-skinny QR -- one chain per warp and strip, DESIGN.md section 8 -- assumes eight chains per SM
// run as fast as one: ~250 issued instructions per 470-cycle reflector step and two chains per scheduler.)
namespace qrb200 {
constexpr int CHAIN_FLOATS = TB * 33 + 2 * TB + 2 * TB * TB + 24;       // Rs, xs, tile, V2 out (+ pad to 16 bytes)
__global__ void __launch_bounds__(512) panel_chains_kernel(const float *tile_in, float *sink, long long *cycles, int reps)
{
    extern __shared__ __align__(16) float csm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *Rs = csm + (size_t)warp * CHAIN_FLOATS;
    float *xs = Rs + TB * 33 + 8;                                        // 16-byte aligned: (33 * 32 + 8) * 4 = 4256
    float *ts = xs + 2 * TB;
    float *ts2 = ts + TB * TB;
    float a[TB], b[TB];
    load_col_gmem(a, tile_in + lane * TB);
    store_col_smem(ts + lane * TB, a);
#pragma unroll
    for (int r = 0; r < TB; ++r) Rs[r * 33 + lane] = (r <= lane) ? a[r] + (r == lane ? 8.0f : 0.f) : 0.f;   // a triangle with a healthy diagonal
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        load_col_smem(b, ts + lane * TB);
        acc += tsqrt_warp_v<false>(b, Rs, xs, ts2, lane);
    }
    const long long t1 = clock64();
    if (lane == 0) { cycles[warp] = (t1 - t0) / reps; sink[warp] = acc; }
}
}  // namespace qrb200

extern "C" int b200qr_probe_panel_chains(long long *cycles, int reps, int nwarps)
{
    if (!cycles || reps <= 0 || nwarps < 1 || nwarps > 16) return B200QR_EINVAL;
    float h[32 * 32];
    for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 37 % 201) - 100) / 100.0f;
    float *din = nullptr, *dsink = nullptr;
    long long *dc = nullptr;
    if (cudaMalloc(&din, 4096) || cudaMalloc(&dsink, 64 * 4) || cudaMalloc(&dc, 16 * 8)) return B200QR_ECUDA;
    cudaMemcpy(din, h, 4096, cudaMemcpyHostToDevice);
    const int smem = nwarps * qrb200::CHAIN_FLOATS * (int)sizeof(float);
    int rc = B200QR_OK;
    if (cudaFuncSetAttribute(qrb200::panel_chains_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) rc = B200QR_ECUDA;
    if (rc == B200QR_OK) {
        qrb200::panel_chains_kernel<<<1, nwarps * 32, smem>>>(din, dsink, dc, reps);
        if (cudaDeviceSynchronize() != cudaSuccess) rc = B200QR_ECUDA;
        else cudaMemcpy(cycles, dc, sizeof(long long) * nwarps, cudaMemcpyDeviceToHost);
    }
    cudaFree(din); cudaFree(dsink); cudaFree(dc);
    return rc;
}
