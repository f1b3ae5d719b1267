// umma_probe.cu -- stand-alone check of the tcgen05 path the trailing update is being moved to:
//   Z (128 x 32) = A^T V   with A a 32 x 128 block of four 32x32 tiles (one tile row of a 128-column
//   group) and V a 32x32 reflector tile, both column-major, brought in by TMA with 128B swizzle,
//   split into 3xTF32 operands in shared memory, multiplied by tcgen05.mma kind::tf32 into TMEM and
//   read back with tcgen05.ld -- one matrix column per thread, which is the thread-per-column layout
//   the SIMT part of the update works in.  mode 1 additionally computes  A^T - W V^T  (the second
//   product of the compact-WY update) from a 128 x 32 W.  Both B operands are K-major: V for the first
//   product, a transposed copy V^T (made during the hi/lo split) for the second -- tcgen05.mma returned
//   zeros on B200 when the same tile was described as an MN-major tf32 operand.
#include "../qr_umma.cuh"
#include "../../../include/qrb200.h"
#include "../../../tools/qrb200_probe.h"

namespace qrb200 {

struct ProbeSmem {
    float a_hi[4 * TB * TB];     // 4 tiles, 16 KB, 1024-aligned (operand A of product 1: M = 128 columns, K = 32 rows)
    float a_lo[4 * TB * TB];
    float v_hi[TB * TB];
    float v_lo[TB * TB];
    float vt_hi[TB * TB];        // V^T: line r = tile row r, 32 reflector entries (operand B of product 2)
    float vt_lo[TB * TB];
    float w_hi[4 * TB * TB];     // operand A of product 2: M = 128 columns, K = 32 reflectors (line m = 128 B)
    float w_lo[4 * TB * TB];
    uint64_t bar_tma, bar_mma;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                         const __grid_constant__ CUtensorMap mapV, const float *W,
                                                         float *Z, float *D2, int mode)
{
    extern __shared__ __align__(1024) unsigned char psm_raw[];
    ProbeSmem &sm = *reinterpret_cast<ProbeSmem *>(psm_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sm.bar_tma, 1);
        mbar_init(&sm.bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<128>(&sm.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(&sm.bar_tma, 5 * TB * TB * 4);
        for (int t = 0; t < 4; ++t) tma_load_2d(sm.a_hi + t * TB * TB, &mapA, 0, t * TB, &sm.bar_tma);
        tma_load_2d(sm.v_hi, &mapV, 0, 0, &sm.bar_tma);
    }
    mbar_wait(&sm.bar_tma, 0);

    // keep the original fp32 column for the final subtraction (mode 1), then split into hi / lo
    float orig[TB];
    load_col_sw(orig, sm.a_hi + warp * TB * TB, tid & 31);
    split_line_sw(sm.a_hi, sm.a_lo, tid);                     // line tid = matrix column tid (tile tid/32, line tid%32)
    if (tid < TB) {                                            // reflector column tid: split, and scatter the transpose
        float v[TB];
        load_col_sw(v, sm.v_hi, tid);
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const float h = tf32_hi(v[r]);
            const int off = r * 32 + ((((tid >> 2) ^ (r & 7)) << 2) | (tid & 3));   // line r, element tid, swizzled
            sm.vt_hi[off] = h;
            sm.vt_lo[off] = v[r] - h;
        }
        split_line_sw(sm.v_hi, sm.v_lo, tid);
    }
    if (mode & 1) {
        // W row m (32 reflector coefficients of matrix column m) -> line m of w_hi / w_lo, swizzled like a TMA tile
        float w[TB];
#pragma unroll
        for (int c = 0; c < TB; ++c) w[c] = W[tid * TB + c];
        store_col_sw(sm.w_hi + warp * TB * TB, tid & 31, w);
        split_line_sw(sm.w_hi, sm.w_lo, tid);
    }
    if (mode & 4) {
        // A operand of product 1 in tensor memory: this thread's column, hi in TMEM columns 64..95, lo in 96..127
        float h[TB], l[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) { h[r] = tf32_hi(orig[r]); l[r] = orig[r] - h[r]; }
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 64, h);
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 96, l);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
        tc_fence_after();
        // product 1: Z = A^T V.  A operand: M = 128 lines (matrix columns) x K = 32, K-major; B operand: N = 32
        // lines (reflectors) x K = 32, K-major.  Four K = 8 steps, 32 bytes apart inside the swizzled line.
        const uint32_t idesc = umma_idesc_tf32(128, 32, false, false);
        const float *As[3] = {sm.a_hi, sm.a_hi, sm.a_lo};
        const float *Bs[3] = {sm.v_hi, sm.v_lo, sm.v_hi};
        bool acc = false;
        const uint32_t At[3] = {tmem + 64, tmem + 64, tmem + 96};
        for (int pass = 0; pass < 3; ++pass)
            for (int kk = 0; kk < 4; ++kk) {
                if (mode & 4) umma_tf32_ts(tmem, At[pass] + kk * 8, umma_desc_sw128(Bs[pass], kk * 32), idesc, acc);
                else umma_tf32(tmem, umma_desc_sw128(As[pass], kk * 32), umma_desc_sw128(Bs[pass], kk * 32), idesc, acc);
                acc = true;
            }
        if (mode & 1) {
            // product 2: P = W V^T into TMEM columns 32..63.  A operand: W, M = 128 lines x K = 32 (reflector index),
            // K-major.  B operand: V^T as N = 32 lines (tile rows) x K = 32 (reflectors), K-major.
            const uint32_t idesc2 = umma_idesc_tf32(128, 32, false, false);
            const float *A2[3] = {sm.w_hi, sm.w_hi, sm.w_lo};
            const float *B2[3] = {sm.vt_hi, sm.vt_lo, sm.vt_hi};
            bool acc2 = false;
            if (mode & 8) {
                // B operand = the reflector tile itself described as MN-major (N = tile rows contiguous in a 128-byte line,
                // K = reflector index: one 1024-byte swizzle atom per K = 8 step).  (mode >> 4) & 7 picks the LBO / SBO variant.
                const uint32_t idesc_mn = umma_idesc_tf32(128, 32, false, true);
                const float *BM[3] = {sm.v_hi, sm.v_lo, sm.v_hi};
                const int var = (mode >> 4) & 7;
                const uint64_t lbo = (var == 0) ? 1 : ((var == 3) ? (4096 >> 4) : (1024 >> 4));
                const uint64_t sbo = (var == 2) ? 1 : (1024 >> 4);
                for (int pass = 0; pass < 3; ++pass)
                    for (int kk = 0; kk < 4; ++kk) {
                        uint64_t d = 0;
                        d |= (uint64_t)(((smem_u32(BM[pass]) + kk * 1024) >> 4) & 0x3fff);
                        d |= lbo << 16;
                        d |= sbo << 32;
                        d |= (uint64_t)1 << 46;
                        d |= (uint64_t)2 << 61;
                        umma_tf32(tmem + 32, umma_desc_sw128(A2[pass], kk * 32), d, idesc_mn, acc2);
                        acc2 = true;
                    }
            } else
            for (int pass = 0; pass < 3; ++pass)
                for (int kk = 0; kk < 4; ++kk) {
                    umma_tf32(tmem + 32, umma_desc_sw128(A2[pass], kk * 32), umma_desc_sw128(B2[pass], kk * 32), idesc2, acc2);
                    acc2 = true;
                }
        }
        umma_commit(&sm.bar_mma);
    }
    mbar_wait(&sm.bar_mma, 0);
    tc_fence_after();

    float z[TB];
    tmem_ld_32x32(z, tmem + ((uint32_t)(warp * 32) << 16));
#pragma unroll
    for (int c = 0; c < TB; ++c) Z[tid * TB + c] = z[c];
    if (mode & 1) {
        float pr[TB];
        tmem_ld_32x32(pr, tmem + ((uint32_t)(warp * 32) << 16) + 32);
#pragma unroll
        for (int r = 0; r < TB; ++r) D2[tid * TB + r] = (mode & 2) ? pr[r] : orig[r] - pr[r];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

}  // namespace qrb200

using namespace qrb200;

typedef CUresult (*probe_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int probe_map(CUtensorMap *map, float *d, int rows, int cols)
{
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return B200QR_ECUDA;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)rows * sizeof(float)};
    cuuint32_t box[2] = {TB, TB}, estr[2] = {1, 1};
    CUresult r = ((probe_encode_fn)ptr)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? B200QR_OK : B200QR_ECUDA;
}

// Host buffers: hA 32 x 128 column-major (ld 32), hV 32 x 32 column-major, hW 128 x 32 row-major (may be NULL);
// hZ receives A^T V (128 x 32 row-major); hD2 (if hW) receives A^T - W V^T (128 x 32 row-major).
extern "C" int b200qr_probe_umma_mode(const float *hA, const float *hV, const float *hW, float *hZ, float *hD2, int mode);
extern "C" int b200qr_probe_umma(const float *hA, const float *hV, const float *hW, float *hZ, float *hD2)
{
    return b200qr_probe_umma_mode(hA, hV, hW, hZ, hD2, hW ? 1 : 0);
}
extern "C" int b200qr_probe_umma_mode(const float *hA, const float *hV, const float *hW, float *hZ, float *hD2, int mode)
{
    if (!hA || !hV || !hZ || (hW && !hD2)) return B200QR_EINVAL;
    float *dA = nullptr, *dV = nullptr, *dW = nullptr, *dZ = nullptr, *dD = nullptr;
    if (cudaMalloc(&dA, 32 * 128 * 4) || cudaMalloc(&dV, 32 * 32 * 4) || cudaMalloc(&dW, 128 * 32 * 4) ||
        cudaMalloc(&dZ, 128 * 32 * 4) || cudaMalloc(&dD, 128 * 32 * 4))
        return B200QR_ECUDA;
    cudaMemcpy(dA, hA, 32 * 128 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV, 32 * 32 * 4, cudaMemcpyHostToDevice);
    if (hW) cudaMemcpy(dW, hW, 128 * 32 * 4, cudaMemcpyHostToDevice);
    CUtensorMap mapA, mapV;
    int rc = probe_map(&mapA, dA, 32, 128);
    if (rc == B200QR_OK) rc = probe_map(&mapV, dV, 32, 32);
    if (rc == B200QR_OK) {
        const int smem = (int)sizeof(ProbeSmem) + 1024;
        cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        umma_probe_kernel<<<1, 128, smem>>>(mapA, mapV, dW, dZ, dD, mode);
        if (cudaDeviceSynchronize() != cudaSuccess) rc = B200QR_ECUDA;
    }
    if (rc == B200QR_OK) {
        cudaMemcpy(hZ, dZ, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        if (hW) cudaMemcpy(hD2, dD, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dA); cudaFree(dV); cudaFree(dW); cudaFree(dZ); cudaFree(dD);
    return rc;
}

// ---- issue-rate probe: cycles per tcgen05.mma (kind::tf32, M = 128, A in tensor memory, K = 8) as a function of N,
// for one accumulator (dependent accumulation) and for NACC accumulators used round-robin.  One CTA, one issuing thread.
namespace qrb200 {
__global__ void __launch_bounds__(128) umma_rate_kernel(long long *out, int reps)
{
    extern __shared__ __align__(1024) unsigned char rsm[];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<float *>(rsm)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __reduce_max_sync(FULL, slot);
    {   // define the A operand region (columns 448..511)
        float z[TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) z[i] = 0.f;
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 448, z);
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 480, z);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t bdesc = ((smem_u32(rsm) >> 4) & 0x3fff) | (1u << 16);
    uint32_t phase = 0;
    int slot_out = 0;
    const int Ns[4] = {32, 64, 128, 256};
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
        for (int nacc = 1; nacc <= 2; ++nacc) {
            if (Ns[ni] * nacc > 384) { if (tid == 0) out[slot_out] = 0; ++slot_out; continue; }
            const uint32_t idesc = umma_idesc_tf32(128, Ns[ni], false, false);
            long long t0 = 0;
            if (tid == 0) {
                t0 = clock64();
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_tf32_ts(tmem + (uint32_t)(((r * 4 + kk) % nacc) * Ns[ni]), tmem + 448 + 8 * kk, umma_desc_from(bdesc + 2 * kk), idesc,
                                     true);
                }
                umma_commit(&bar_done);
            }
            mbar_wait(&bar_done, phase);
            phase ^= 1;
            if (tid == 0) out[slot_out] = clock64() - t0;
            ++slot_out;
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace qrb200

// Do two issuing threads (different warps, different accumulators) double the tcgen05.mma rate, or is the ~46-cycle
// floor per instruction a property of the tensor pipe?  out2[0]: cycles for one thread issuing reps*4 MMAs (N = 32);
// out2[1]: cycles until BOTH of two threads have finished reps*4 MMAs each.
namespace qrb200 {
__global__ void __launch_bounds__(128) umma_rate2_kernel(long long *out, int reps)
{
    extern __shared__ __align__(1024) unsigned char rsm2[];
    __shared__ uint64_t bar_done[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<float *>(rsm2)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __reduce_max_sync(FULL, slot);
    {
        float z[TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) z[i] = 0.f;
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 448, z);
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 480, z);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t bdesc = ((smem_u32(rsm2) >> 4) & 0x3fff) | (1u << 16);
    const uint32_t idesc = umma_idesc_tf32(128, 32, false, false);
    for (int nthr = 1; nthr <= 2; ++nthr) {
        __syncthreads();
        const long long t0 = clock64();
        if ((tid & 31) == 0 && warp < nthr) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_tf32_ts(tmem + (uint32_t)(warp * 64), tmem + 448 + 8 * kk, umma_desc_from(bdesc + 2 * kk), idesc, true);
            }
            umma_commit(&bar_done[warp]);
        }
        for (int w = 0; w < nthr; ++w) mbar_wait(&bar_done[w], (uint32_t)(nthr - 1 - (w == 1 ? 1 : 0)) & 1);
        if (tid == 0) out[nthr - 1] = clock64() - t0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace qrb200

extern "C" int b200qr_probe_umma_rate2(long long *out2, int reps)
{
    if (!out2 || reps <= 0) return B200QR_EINVAL;
    long long *d = nullptr;
    if (cudaMalloc(&d, 2 * sizeof(long long)) != cudaSuccess) return B200QR_ECUDA;
    cudaFuncSetAttribute(umma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 33 * 1024);
    umma_rate2_kernel<<<1, 128, 33 * 1024>>>(d, reps);
    int rc = cudaDeviceSynchronize() == cudaSuccess ? B200QR_OK : B200QR_ECUDA;
    if (rc == B200QR_OK) cudaMemcpy(out2, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return rc;
}

// out8: cycles for reps*4 MMAs: N = 32 (1 acc, 2 acc), 64 (1, 2), 128 (1, 2), 256 (1, -)
extern "C" int b200qr_probe_umma_rate(long long *out8, int reps)
{
    if (!out8 || reps <= 0) return B200QR_EINVAL;
    long long *d = nullptr;
    if (cudaMalloc(&d, 8 * sizeof(long long)) != cudaSuccess) return B200QR_ECUDA;
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 33 * 1024);
    umma_rate_kernel<<<1, 128, 33 * 1024>>>(d, reps);
    int rc = cudaDeviceSynchronize() == cudaSuccess ? B200QR_OK : B200QR_ECUDA;
    if (rc == B200QR_OK) cudaMemcpy(out8, d, 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return rc;
}
