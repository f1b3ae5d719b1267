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

