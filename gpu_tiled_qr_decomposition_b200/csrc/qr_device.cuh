// qr_device.cuh -- device building blocks of the B200 tiled QR (sm_100a).
//
// Thread mapping used everywhere: ONE THREAD OWNS ONE MATRIX COLUMN of the 32-row tile
// it works on (32 floats in registers).  A warp therefore owns one 32x32 tile.
//   * panel kernels (GEQRT / TSQRT): the Householder norm and the trailing dot products
//     are thread-local FMA chains; the reflector is broadcast with warp shuffles;
//   * update kernels (UNMQR / TSMQR): no cross-thread communication at all -- the
//     reflector block V (and tau) sits in shared memory and is read with 128-bit
//     broadcast loads, the column being updated never leaves the register file.
// Tiles move HBM <-> shared memory with bulk async copies (cp.async.bulk, the TMA
// engine) completing on mbarriers; each thread moves its own 128-byte column.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qrb200 {

constexpr int TB = 32;               // tile size (reference: gpucalc.cu:1582, qrdecomp.c:506)
constexpr int COL_STRIDE = 36;       // floats between columns in smem staging (144 B: conflict-free LDS.128)
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// generic-proxy <-> async-proxy ordering (flags are read/written through the generic proxy,
// tiles through the async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void load_col_smem(float (&x)[TB], const float *col)
{
#pragma unroll
    for (int r = 0; r < TB; r += 4) {
        float4 t = *reinterpret_cast<const float4 *>(col + r);
        x[r] = t.x; x[r + 1] = t.y; x[r + 2] = t.z; x[r + 3] = t.w;
    }
}
__device__ __forceinline__ void store_col_smem(float *col, const float (&x)[TB])
{
#pragma unroll
    for (int r = 0; r < TB; r += 4)
        *reinterpret_cast<float4 *>(col + r) = make_float4(x[r], x[r + 1], x[r + 2], x[r + 3]);
}
__device__ __forceinline__ void load_col_gmem(float (&x)[TB], const float *col)
{
#pragma unroll
    for (int r = 0; r < TB; r += 4) {
        float4 t = __ldcg(reinterpret_cast<const float4 *>(col + r));
        x[r] = t.x; x[r + 1] = t.y; x[r + 2] = t.z; x[r + 3] = t.w;
    }
}
__device__ __forceinline__ void store_col_gmem(float *col, const float (&x)[TB])
{
#pragma unroll
    for (int r = 0; r < TB; r += 4)
        *reinterpret_cast<float4 *>(col + r) = make_float4(x[r], x[r + 1], x[r + 2], x[r + 3]);
}

// ------------------------------------------------------------------ a1: GEQRT ("QRS")
// Householder QR of one 32x32 tile held by a warp, lane = column (reference semantics:
// qrdecomp.c:532-557 SGEQRF, 1201-1227 calcvkSingle, 489-530 updateSingleQ_WY).
// On exit a[] holds R on/above the diagonal and the reflector below it; the return value
// of lane c is tau_c = 2/(1+|v_ess|^2).
__device__ __forceinline__ float geqrt_warp(float (&a)[TB], int lane)
{
    float my_tau = 0.f;
#pragma unroll 1
    for (int c = 0; c < TB; ++c) {
        // |x|^2 of this lane's column from row c down, and its head element x_c
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, head = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            float y0 = (r + 0 >= c) ? a[r + 0] : 0.f;
            float y1 = (r + 1 >= c) ? a[r + 1] : 0.f;
            float y2 = (r + 2 >= c) ? a[r + 2] : 0.f;
            float y3 = (r + 3 >= c) ? a[r + 3] : 0.f;
            s0 = fmaf(y0, y0, s0); s1 = fmaf(y1, y1, s1);
            s2 = fmaf(y2, y2, s2); s3 = fmaf(y3, y3, s3);
            head = (r + 0 == c) ? a[r + 0] : head;
            head = (r + 1 == c) ? a[r + 1] : head;
            head = (r + 2 == c) ? a[r + 2] : head;
            head = (r + 3 == c) ? a[r + 3] : head;
        }
        float ssq = (s0 + s1) + (s2 + s3);
        // only column c defines the reflector
        ssq = __shfl_sync(FULL, ssq, c);
        head = __shfl_sync(FULL, head, c);
        float nrm = sqrtf(ssq);
        float piv = head + (head >= 0.f ? nrm : -nrm);      // x_c + sign(x_c)|x|, sign(0)=+1 (qrdecomp.c:1208)
        float inv = (nrm != 0.f) ? 1.0f / piv : 1.0f;       // |x|==0: left unscaled (qrdecomp.c:1219)
        // v = [0 .. 0, 1, x_{c+1..}/piv], broadcast from lane c
        float v[TB];
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            float xr = __shfl_sync(FULL, a[r], c);
            v[r] = (r > c) ? xr * inv : ((r == c) ? 1.0f : 0.f);
        }
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            q0 = fmaf(v[r], v[r], q0); q1 = fmaf(v[r + 1], v[r + 1], q1);
            q2 = fmaf(v[r + 2], v[r + 2], q2); q3 = fmaf(v[r + 3], v[r + 3], q3);
        }
        float scale = -2.0f / ((q0 + q1) + (q2 + q3));       // -tau_c
        // A(:, j) += scale * v * (v' A(:, j)) for j >= c
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            d0 = fmaf(a[r], v[r], d0); d1 = fmaf(a[r + 1], v[r + 1], d1);
            d2 = fmaf(a[r + 2], v[r + 2], d2); d3 = fmaf(a[r + 3], v[r + 3], d3);
        }
        float f = (lane >= c) ? scale * ((d0 + d1) + (d2 + d3)) : 0.f;
#pragma unroll
        for (int r = 0; r < TB; ++r) a[r] = fmaf(f, v[r], a[r]);
        if (lane == c) {
#pragma unroll
            for (int r = 0; r < TB; ++r) a[r] = (r > c) ? v[r] : a[r];
            my_tau = -scale;
        }
    }
    return my_tau;
}

// ------------------------------------------------------------------ a2: TSQRT ("QRD")
// QR of [R; B] with R upper triangular (kept in shared memory, Rs[row*33 + col]) and B a
// full 32x32 tile held by the warp, lane = column (reference: qrdecomp.c:689-721 STSQRF,
// 1243-1272 calcvkDouble, 647-687 updateDoubleQ_WY).  On exit b[] holds V2, Rs the new R,
// lane c returns tau_c.
__device__ __forceinline__ float tsqrt_warp(float (&b)[TB], float *Rs, int lane)
{
    float my_tau = 0.f;
#pragma unroll 1
    for (int c = 0; c < TB; ++c) {
        float rc = Rs[c * 33 + lane];                         // R(c, lane)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            s0 = fmaf(b[r], b[r], s0); s1 = fmaf(b[r + 1], b[r + 1], s1);
            s2 = fmaf(b[r + 2], b[r + 2], s2); s3 = fmaf(b[r + 3], b[r + 3], s3);
        }
        float ssq = (s0 + s1) + (s2 + s3);
        float diag = __shfl_sync(FULL, rc, c);
        ssq = __shfl_sync(FULL, ssq, c);
        float nrm = sqrtf(fmaf(diag, diag, ssq));
        float piv = diag + (diag >= 0.f ? nrm : -nrm);
        float inv = (nrm != 0.f) ? 1.0f / piv : 1.0f;
        float v[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) v[r] = __shfl_sync(FULL, b[r], c) * inv;
        float q0 = 1.0f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            q0 = fmaf(v[r], v[r], q0); q1 = fmaf(v[r + 1], v[r + 1], q1);
            q2 = fmaf(v[r + 2], v[r + 2], q2); q3 = fmaf(v[r + 3], v[r + 3], q3);
        }
        float t = 2.0f / ((q0 + q1) + (q2 + q3));
        float d0 = rc, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            d0 = fmaf(b[r], v[r], d0); d1 = fmaf(b[r + 1], v[r + 1], d1);
            d2 = fmaf(b[r + 2], v[r + 2], d2); d3 = fmaf(b[r + 3], v[r + 3], d3);
        }
        float beta = (lane >= c) ? t * ((d0 + d1) + (d2 + d3)) : 0.f;
        Rs[c * 33 + lane] = rc - beta;
#pragma unroll
        for (int r = 0; r < TB; ++r) b[r] = fmaf(-beta, v[r], b[r]);
        if (lane == c) {
#pragma unroll
            for (int r = 0; r < TB; ++r) b[r] = v[r];
            my_tau = t;
        }
    }
    return my_tau;
}

// ------------------------------------------------------------------ a3/a4: UNMQR / TSMQR
// Apply the 32 reflectors of one block to the column owned by this thread.
//   vs  : reflector block in shared memory, reflector c at vs[c*32 .. c*32+31]
//   ts  : tau_0..tau_31 in shared memory
//   HAS_TOP = true  (TSMQR / "DAPP", qrdecomp.c:723-763): v_c = [e_c ; V2(:,c)], the top
//             part of the column is top[c*top_stride] in shared memory, b[] the bottom;
//   HAS_TOP = false (UNMQR / "SAPP", qrdecomp.c:620-644): vs already holds the explicit
//             unit-lower-triangular V (zeros above, ones on the diagonal), b[] is the column.
// Rank-1 updates applied one after another, exactly the reference's operation sequence:
// beta = tau_c (top_c + v_c' b); top_c -= beta; b -= beta v_c.
template <bool HAS_TOP>
__device__ __forceinline__ void apply_block_reflectors(float (&b)[TB], float *top, int top_stride,
                                                       const float *vs, const float *ts)
{
#pragma unroll 2
    for (int c = 0; c < TB; ++c) {
        float v[TB];
        load_col_smem(v, vs + c * TB);
        float tc = ts[c];
        float a0 = 0.f;
        if (HAS_TOP) a0 = top[c * top_stride];
        float d0 = a0, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int r = 0; r < TB; r += 4) {
            d0 = fmaf(b[r], v[r], d0); d1 = fmaf(b[r + 1], v[r + 1], d1);
            d2 = fmaf(b[r + 2], v[r + 2], d2); d3 = fmaf(b[r + 3], v[r + 3], d3);
        }
        float beta = tc * ((d0 + d1) + (d2 + d3));
        if (HAS_TOP) top[c * top_stride] = a0 - beta;
#pragma unroll
        for (int r = 0; r < TB; ++r) b[r] = fmaf(-beta, v[r], b[r]);
    }
}

// Turn a raw diagonal tile (R above, GEQRT reflectors below) staged in shared memory into the
// explicit unit-lower-triangular V that apply_block_reflectors<false> expects.
__device__ __forceinline__ void make_unit_lower(float *vs, int tid, int nthreads)
{
    for (int e = tid; e < TB * TB; e += nthreads) {
        int c = e >> 5, r = e & 31;
        if (r < c) vs[e] = 0.f;
        else if (r == c) vs[e] = 1.0f;
    }
}

}  // namespace qrb200
