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
#include <cuda.h>
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
// plain L2 poll (LDG.STRONG.GPU, no L1 invalidate); pair with one __threadfence() after the loop
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// system-scope variants: the word lives in (or is read from) a peer GPU's memory over NVLink
__device__ __forceinline__ int ld_relaxed_sys(const int *p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(int *p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
// non-blocking test (try_wait may suspend the thread for a while before it reports failure)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
// TMA: 2-D tiled tensor copies (SASS: UTMALDG / UTMASTG).  c0 = row (inner), c1 = column.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src_smem)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
// A 32x32 fp32 tile landed by TMA with CU_TENSOR_MAP_SWIZZLE_128B: matrix column c is the 128-byte
// line c of the (1024-byte aligned) tile, and its 16-byte chunk i sits at chunk slot i ^ (c & 7).
// A thread that owns column `lane` therefore reads 8 x LDS.128 with no bank conflicts.
__device__ __forceinline__ void load_col_sw(float (&x)[TB], const float *tile, int lane)
{
    const char *base = reinterpret_cast<const char *>(tile) + lane * 128;
    const int sw = (lane & 7) << 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 t = *reinterpret_cast<const float4 *>(base + ((i << 4) ^ sw));
        x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_col_sw(float *tile, int lane, const float (&x)[TB])
{
    char *base = reinterpret_cast<char *>(tile) + lane * 128;
    const int sw = (lane & 7) << 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4 *>(base + ((i << 4) ^ sw)) =
            make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
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

// ------------------------------------------------------------------ panel kernels: shared pieces
// Householder scalars from the head element x0 and the sum of squares of the rest, restating
// calcvkSingle / calcvkDouble + the tau of updateSingleQ_WY / updateDoubleQ_WY
// (qrdecomp.c:1201-1272, 498-504, 660-664):
//   nrm = sqrt(x0^2 + ssq);  piv = x0 + sign(x0) nrm (sign(0) = +1);  v = x / piv;
//   tau = 2 / (1 + |v_ess|^2)  ==  |piv| / nrm  ==  1 + |x0| / nrm      (exact identity),
// so one reciprocal square root gives nrm and tau, and 1/piv does not wait for a second
// division.  nrm == 0 leaves v unscaled and tau = 2, like the reference (qrdecomp.c:1219).
struct HouseScalars { float inv, tau; };
__device__ __forceinline__ HouseScalars house_scalars(float x0, float ssq)
{
    const float nsq = fmaf(x0, x0, ssq);
    float rn, r0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(nsq));
    rn = rn * fmaf(-0.5f * nsq * rn, rn, 1.5f);            // one Newton step: ~1 ulp
    const float nrm = nsq * rn;
    const float piv = x0 + (x0 >= 0.f ? nrm : -nrm);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(piv));
    r0 = fmaf(r0, fmaf(-piv, r0, 1.0f), r0);               // one Newton step
    HouseScalars h;
    const bool degenerate = !(nsq > 1e-36f);               // zero (or flushed) column: v stays unscaled, tau = 2
    h.inv = degenerate ? 1.0f : r0;
    h.tau = degenerate ? 2.0f : fmaf(fabsf(x0), rn, 1.0f);
    return h;
}
// 8 independent accumulators: 4-deep FMA chains instead of 32-deep
__device__ __forceinline__ float dot32(const float (&a)[TB], const float (&x)[TB])
{
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = a[i] * x[i];
#pragma unroll
    for (int r = 8; r < TB; r += 8)
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = fmaf(a[r + i], x[r + i], s[i]);
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
// one lane's 32 floats -> every lane, through a 128-byte shared-memory slot
__device__ __forceinline__ void bcast_col(float (&x)[TB], const float (&mine)[TB], float *slot, bool owner)
{
    if (owner) store_col_smem(slot, mine);
    __syncwarp();
    load_col_smem(x, slot);
}

// ------------------------------------------------------------------ a2: TSQRT ("QRD")
// QR of [R; B] with R upper triangular (kept in shared memory, Rs[row*33 + col]) and B a
// full 32x32 tile held by the warp, lane = column (reference: qrdecomp.c:689-721 STSQRF,
// 1243-1272 calcvkDouble, 647-687 updateDoubleQ_WY).  On exit b[] holds V2, Rs the new R,
// lane c returns tau_c.  The serial chain per column is kept short: the owner lane's sum of
// squares is ready from the previous update, the dot products use the unscaled column x (the
// scale 1/piv is applied to the scalar), so they overlap the rsqrt / reciprocal.
struct NoMid { __device__ __forceinline__ void operator()() const {} };
// `mid` runs once, a quarter of the way through the tile: the persistent kernel uses it to publish
// the PREVIOUS tile of the chain (whose TMA store has long completed by then) off the critical path.
// The reflector block V2 is assembled in the shared-memory tile `vout` (lane r writes V2(r,c) at
// step c, one multiply per lane instead of 32 by the owner lane): element (r,c) at vout[c*32 + r],
// or in the 128B-swizzled TMA layout when SWZ.  b[] is clobbered.
template <bool SWZ, class Mid = NoMid>
__device__ __forceinline__ float tsqrt_warp_v(float (&b)[TB], float *Rs, float *xs, float *vout, int lane, Mid mid = Mid())
{
    float my_tau = 0.f;
    float ssq = dot32(b, b);
    float *vdst = SWZ ? reinterpret_cast<float *>(reinterpret_cast<char *>(vout) + (lane & 3) * 4) : vout + lane;
#pragma unroll 1
    for (int c = 0; c < TB; ++c) {
        if (c == 8) mid();
        float *slot = xs + (c & 1) * TB;
        if (lane == c) store_col_smem(slot, b);
        const float ssq_c = __shfl_sync(FULL, ssq, c);
        const float rc = Rs[c * 33 + lane];                   // R(c, lane)
        const float diag = Rs[c * 33 + c];                    // R(c, c): broadcast read
        __syncwarp();
        float x[TB];
        load_col_smem(x, slot);                               // x = B(:, c)
        const HouseScalars h = house_scalars(diag, ssq_c);
        // Every lane forms the dot product, unconditionally: with no branch in the loop body the scheduler runs the
        // rsqrt / rcp chain of the Householder scalars underneath the shared-memory broadcast and the dot product
        // (round 1 had a lane >= c branch here: 15.3k cycles per tile, now 13.0k; tools/probe_panel.py).
        float d = dot32(b, x);
        asm volatile("" : "+f"(d));
        // beta = tau (R(c,j) + v' B(:,j)),  v = x / piv
        const float beta = (lane >= c) ? h.tau * fmaf(h.inv, d, rc) : 0.f;
        Rs[c * 33 + lane] = rc - beta;
        const float g = (lane > c) ? -beta * h.inv : 0.f;
#pragma unroll
        for (int r = 0; r < TB; ++r) b[r] = fmaf(g, x[r], b[r]);
        ssq = dot32(b, b);                                    // needed from lane c+1 next step
        const float vr = slot[lane] * h.inv;                  // V2(lane, c)
        if (SWZ) *reinterpret_cast<float *>(reinterpret_cast<char *>(vdst) + c * 128 + ((((lane >> 2) ^ (c & 7))) << 4)) = vr;
        else vdst[c * TB] = vr;
        if (lane == c) my_tau = h.tau;
    }
    __syncwarp();
    return my_tau;
}
// register-output flavour (single-tile kernels, batched kernel): V2 comes back in b[]
template <class Mid = NoMid>
__device__ __forceinline__ float tsqrt_warp(float (&b)[TB], float *Rs, float *xs, float *vtmp, int lane, Mid mid = Mid())
{
    const float t = tsqrt_warp_v<false>(b, Rs, xs, vtmp, lane, mid);
    load_col_smem(b, vtmp + lane * TB);
    __syncwarp();
    return t;
}

// ------------------------------------------------------------------ a1: GEQRT ("QRS")
// Householder QR of one 32x32 tile held by a warp, lane = column (reference semantics: qrdecomp.c:532-557 SGEQRF,
// 1201-1227 calcvkSingle, 489-530 updateSingleQ_WY).  On exit a[] holds R on/above the diagonal and the reflector
// below it; the return value of lane c is tau_c.  xs: 64 floats of shared-memory scratch (16-byte aligned).
// Fully unrolled over the reflector index: every row mask (r > c) is a compile-time fact, so the passes
// over a column shrink with c (16 rows on average instead of 32 with 96 selects per reflector), the dot product
// runs against the UNSCALED column (its scale 1/piv goes onto the scalar, off the rsqrt / rcp chain, as in TSQRT)
// and the sub-column sum of squares the next owner needs comes out of the update pass.  Measured (tools/
// probe_panel.py): 9.9k cycles per tile against 30.9k for round 1's rolled loop with runtime masks.
template <int C>
__device__ __forceinline__ void geqrt_step(float (&a)[TB], float *xs, int lane, float &ssq_next, float &my_tau)
{
    float *slot = xs + (C & 1) * TB;
    constexpr int R0 = C & ~3;                                // first 4-row chunk that holds a row >= C
    if (lane == C) {
#pragma unroll
        for (int r = R0; r < TB; r += 4) *reinterpret_cast<float4 *>(slot + r) = make_float4(a[r], a[r + 1], a[r + 2], a[r + 3]);
    }
    const float ssq_c = __shfl_sync(FULL, ssq_next, C);
    __syncwarp();
    float x[TB];
#pragma unroll
    for (int r = R0; r < TB; r += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(slot + r);
        x[r] = t.x; x[r + 1] = t.y; x[r + 2] = t.z; x[r + 3] = t.w;
    }
    const HouseScalars h = house_scalars(x[C], ssq_c);
    // d = sum_{r > C} a[r] x[r]  (four partial sums)
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int r = C + 1; r < TB; ++r) {
        const int s = (r - C - 1) & 3;
        if (s == 0) d0 = fmaf(a[r], x[r], d0);
        else if (s == 1) d1 = fmaf(a[r], x[r], d1);
        else if (s == 2) d2 = fmaf(a[r], x[r], d2);
        else d3 = fmaf(a[r], x[r], d3);
    }
    const float d = (d0 + d1) + (d2 + d3);
    // A(:, j) -= tau v (v' A(:, j)),  v = [1; x / piv]:  f = tau (a_C + d / piv)
    const float f = (lane >= C) ? h.tau * fmaf(h.inv, d, a[C]) : 0.f;
    const float g = (lane > C) ? -f * h.inv : 0.f;
    a[C] -= f;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int r = C + 1; r < TB; ++r) {
        a[r] = fmaf(g, x[r], a[r]);
        if (r > C + 1) {                                       // sum of squares below the NEXT diagonal (lane C + 1 uses it)
            const int s = (r - C - 2) & 3;
            if (s == 0) s0 = fmaf(a[r], a[r], s0);
            else if (s == 1) s1 = fmaf(a[r], a[r], s1);
            else if (s == 2) s2 = fmaf(a[r], a[r], s2);
            else s3 = fmaf(a[r], a[r], s3);
        }
    }
    ssq_next = (s0 + s1) + (s2 + s3);
    if (lane == C) {                                           // the owner keeps the reflector (predicated: one lane)
#pragma unroll
        for (int r = C + 1; r < TB; ++r) a[r] = x[r] * h.inv;
        my_tau = h.tau;
    }
}
template <int C>
__device__ __forceinline__ void geqrt_steps(float (&a)[TB], float *xs, int lane, float &ssq_next, float &my_tau)
{
    if constexpr (C < TB) {
        geqrt_step<C>(a, xs, lane, ssq_next, my_tau);
        geqrt_steps<C + 1>(a, xs, lane, ssq_next, my_tau);
    }
}
__device__ __forceinline__ float geqrt_warp(float (&a)[TB], float *xs, int lane)
{
    float my_tau = 0.f;
    float ssq = 0.f;                                           // sum_{r > 0} a[r]^2 (lane 0's is the one used)
    {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = 1; r < TB; r += 4) {
            s0 = fmaf(a[r], a[r], s0);
            if (r + 1 < TB) s1 = fmaf(a[r + 1], a[r + 1], s1);
            if (r + 2 < TB) s2 = fmaf(a[r + 2], a[r + 2], s2);
            if (r + 3 < TB) s3 = fmaf(a[r + 3], a[r + 3], s3);
        }
        ssq = (s0 + s1) + (s2 + s3);
    }
    geqrt_steps<0>(a, xs, lane, ssq, my_tau);
    __syncwarp();
    return my_tau;
}

// ------------------------------------------------------------------ a1: GEQRT, compact code (tensor-core engine)
// Same kernel as geqrt_warp in a seventh of the code: four phases of eight reflectors.  Inside phase P the rows above 8P are finished and are skipped at compile time,
// the rows from 8P + 8 on are always below the diagonal and need no mask; only the eight rows of the phase carry a
// run-time mask (r > c).  So the passes over a column shrink with the phase (20 rows on average instead of 32 with 96
// selects per reflector), the dot product runs against the UNSCALED column (its scale 1/piv goes onto the scalar, off
// the rsqrt / rcp chain, as in TSQRT) and the sub-column sum of squares the next owner needs comes out of the update
// pass.  Why it exists: geqrt_warp is 72 KB of straight-line code.  In the probe, the SIMT engine and the batched kernel
// that is the fastest form (10.3k cycles per tile; this one: 24k), but inside the tensor-core kernel, whose panel CTAs
// also run the T builders, it missed the instruction cache on every panel -- 22 us per GEQRT at 8192^2 and 48 us at
// 16384^2 in the device trace, against 13 / 17 us for this form (tools/trace_timeline.py n 2).
template <int P8>
__device__ __forceinline__ void geqrt_phase(float (&a)[TB], float *xs, int lane, float &ssq_next, float &my_tau)
{
    constexpr int R0 = 8 * P8;                                 // first row this phase still touches
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
        const int c = R0 + cc;
        float *slot = xs + (c & 1) * TB;
        if (lane == c) {
#pragma unroll
            for (int r = R0; r < TB; r += 4) *reinterpret_cast<float4 *>(slot + r) = make_float4(a[r], a[r + 1], a[r + 2], a[r + 3]);
        }
        const float ssq_c = __shfl_sync(FULL, ssq_next, c);
        __syncwarp();
        float x[TB];
#pragma unroll
        for (int r = R0; r < TB; r += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(slot + r);
            x[r] = t.x; x[r + 1] = t.y; x[r + 2] = t.z; x[r + 3] = t.w;
        }
        const HouseScalars h = house_scalars(slot[c], ssq_c);
        // d = sum_{r > c} a[r] x[r]: the eight rows of the phase masked, the rest below the diagonal for every c of the phase
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, ac = 0.f;
#pragma unroll
        for (int r = R0; r < R0 + 8; r += 2) {
            d0 = r > c ? fmaf(a[r], x[r], d0) : d0;
            d1 = r + 1 > c ? fmaf(a[r + 1], x[r + 1], d1) : d1;
            ac = r == c ? a[r] : (r + 1 == c ? a[r + 1] : ac);
        }
#pragma unroll
        for (int r = R0 + 8; r < TB; r += 4) {
            d0 = fmaf(a[r], x[r], d0); d1 = fmaf(a[r + 1], x[r + 1], d1);
            d2 = fmaf(a[r + 2], x[r + 2], d2); d3 = fmaf(a[r + 3], x[r + 3], d3);
        }
        // A(:, j) -= tau v (v' A(:, j)),  v = [1; x / piv]:  f = tau (a_c + d / piv)
        const float f = (lane >= c) ? h.tau * fmaf(h.inv, (d0 + d1) + (d2 + d3), ac) : 0.f;
        const float g = (lane > c) ? -f * h.inv : 0.f;
        const bool own = lane == c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = R0; r < R0 + 8; r += 2) {
            const float u0 = own ? x[r] * h.inv : fmaf(g, x[r], a[r]);          // the owner keeps the reflector
            const float u1 = own ? x[r + 1] * h.inv : fmaf(g, x[r + 1], a[r + 1]);
            a[r] = r > c ? u0 : (r == c ? a[r] - f : a[r]);
            a[r + 1] = r + 1 > c ? u1 : (r + 1 == c ? a[r + 1] - f : a[r + 1]);
            s2 = r > c + 1 ? fmaf(a[r], a[r], s2) : s2;                          // (lane c + 1 uses it: its column is not the owner's)
            s3 = r + 1 > c + 1 ? fmaf(a[r + 1], a[r + 1], s3) : s3;
        }
#pragma unroll
        for (int r = R0 + 8; r < TB; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[r + u] = own ? x[r + u] * h.inv : fmaf(g, x[r + u], a[r + u]);
            // (row 8P + 8 is the NEXT diagonal row when this is the phase's last reflector: not below it)
            s0 = (r > R0 + 8 || cc < 7) ? fmaf(a[r], a[r], s0) : s0;
            s1 = fmaf(a[r + 1], a[r + 1], s1);
            s2 = fmaf(a[r + 2], a[r + 2], s2); s3 = fmaf(a[r + 3], a[r + 3], s3);
        }
        ssq_next = (s0 + s1) + (s2 + s3);
        if (own) my_tau = h.tau;
    }
}
__device__ __forceinline__ float geqrt_warp_compact(float (&a)[TB], float *xs, int lane)
{
    float my_tau = 0.f;
    float ssq = 0.f;                                           // sum_{r > 0} a[r]^2 (lane 0's is the one used)
    {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = 1; r < TB; r += 4) {
            s0 = fmaf(a[r], a[r], s0);
            if (r + 1 < TB) s1 = fmaf(a[r + 1], a[r + 1], s1);
            if (r + 2 < TB) s2 = fmaf(a[r + 2], a[r + 2], s2);
            if (r + 3 < TB) s3 = fmaf(a[r + 3], a[r + 3], s3);
        }
        ssq = (s0 + s1) + (s2 + s3);
    }
    geqrt_phase<0>(a, xs, lane, ssq, my_tau);
    geqrt_phase<1>(a, xs, lane, ssq, my_tau);
    geqrt_phase<2>(a, xs, lane, ssq, my_tau);
    geqrt_phase<3>(a, xs, lane, ssq, my_tau);
    __syncwarp();
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

// Same operation on NC (1 or 2) columns held by one thread (the persistent kernel's sweep
// consumers): the reflector is read from shared memory once and applied to NC independent
// columns, which also gives the FMA pipe NC independent dependency chains.  top0/top1 point at
// this thread's entries of the resident tile row (element c at top[c*top_rs]).
// HAS_TOP = false takes the raw diagonal tile (R above the diagonal) and substitutes the
// implied zeros / unit diagonal on the fly (UNMQR, qrdecomp.c:620-644).
// VSWZ: vs is a tile in the 128B-swizzled TMA layout (the panel warp applies the block it has just
// assembled in its output tile); otherwise dense, reflector c at vs[c*32 .. c*32+31].
template <int NC, bool HAS_TOP, bool VSWZ = false>
__device__ __forceinline__ void apply_reflectors_nc(float (&b0)[TB], float (&b1)[TB], float *top0, float *top1,
                                                    int top_rs, const float *vs, const float *ts)
{
#pragma unroll 2
    for (int c = 0; c < TB; ++c) {
        float v[TB];
        if (VSWZ) {
            const char *base = reinterpret_cast<const char *>(vs) + c * 128;
            const int sw = (c & 7) << 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 t4 = *reinterpret_cast<const float4 *>(base + ((i << 4) ^ sw));
                v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
            }
        } else {
            load_col_smem(v, vs + c * TB);
        }
        if (!HAS_TOP) {
#pragma unroll
            for (int r = 0; r < TB; ++r) v[r] = (r > c) ? v[r] : ((r == c) ? 1.0f : 0.f);
        }
        const float tc = ts[c];
        {
            float a0 = 0.f;
            if (HAS_TOP) a0 = top0[c * top_rs];
            float d0 = a0, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = 0; r < TB; r += 4) {
                d0 = fmaf(b0[r], v[r], d0); d1 = fmaf(b0[r + 1], v[r + 1], d1);
                d2 = fmaf(b0[r + 2], v[r + 2], d2); d3 = fmaf(b0[r + 3], v[r + 3], d3);
            }
            const float beta = tc * ((d0 + d1) + (d2 + d3));
            if (HAS_TOP) top0[c * top_rs] = a0 - beta;
#pragma unroll
            for (int r = 0; r < TB; ++r) b0[r] = fmaf(-beta, v[r], b0[r]);
        }
        if (NC == 2) {
            float a0 = 0.f;
            if (HAS_TOP) a0 = top1[c * top_rs];
            float d0 = a0, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = 0; r < TB; r += 4) {
                d0 = fmaf(b1[r], v[r], d0); d1 = fmaf(b1[r + 1], v[r + 1], d1);
                d2 = fmaf(b1[r + 2], v[r + 2], d2); d3 = fmaf(b1[r + 3], v[r + 3], d3);
            }
            const float beta = tc * ((d0 + d1) + (d2 + d3));
            if (HAS_TOP) top1[c * top_rs] = a0 - beta;
#pragma unroll
            for (int r = 0; r < TB; ++r) b1[r] = fmaf(-beta, v[r], b1[r]);
        }
    }
}

// The same operation on N columns held by one thread, N a template parameter (round 2: N = 4).  What bounds the update
// on the SIMT pipe is not FFMA issue but the shared-memory pipe: the reflector is the same for every thread, yet each
// thread has to load it -- 8 broadcast LDS.128 per reflector and warp, four MIO cycles each -- so a warp with two
// columns per thread spends 36 MIO cycles per 128 FFMA (ncu: mio_throttle / short_scoreboard on top, FMA pipe 34-42 %).
// Four columns per thread halve the loads per FMA; the columns also give the FMA pipe four independent chains.
template <int N, bool HAS_TOP>
__device__ __forceinline__ void apply_reflectors_cols(float (&b)[N][TB], float *const (&top)[N], int top_rs, const float *vs,
                                                      const float *ts)
{
#pragma unroll 2
    for (int c = 0; c < TB; ++c) {
        float v[TB];
        load_col_smem(v, vs + c * TB);
        if (!HAS_TOP) {
#pragma unroll
            for (int r = 0; r < TB; ++r) v[r] = (r > c) ? v[r] : ((r == c) ? 1.0f : 0.f);
        }
        const float tc = ts[c];
        float a0[N], beta[N];
#pragma unroll
        for (int n = 0; n < N; ++n) {
            a0[n] = HAS_TOP ? top[n][c * top_rs] : 0.f;
            float d0 = a0[n], d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int r = 0; r < TB; r += 4) {
                d0 = fmaf(b[n][r], v[r], d0); d1 = fmaf(b[n][r + 1], v[r + 1], d1);
                d2 = fmaf(b[n][r + 2], v[r + 2], d2); d3 = fmaf(b[n][r + 3], v[r + 3], d3);
            }
            beta[n] = tc * ((d0 + d1) + (d2 + d3));
        }
#pragma unroll
        for (int n = 0; n < N; ++n) {
            if (HAS_TOP) top[n][c * top_rs] = a0[n] - beta[n];
#pragma unroll
            for (int r = 0; r < TB; ++r) b[n][r] = fmaf(-beta[n], v[r], b[n][r]);
        }
    }
}

// ------------------------------------------------------------------ compact-WY T factor of one block
// The tensor-core sweep engine (qr_persistent_tc3.cuh) applies a block of 32 reflectors as
// C -= V (T^T (V^T C)) -- the reference's reflector-by-reflector update (qrdecomp.c:620-644, 736-762)
// in compact-WY form, which the reference itself never forms:
//     T[c][c] = tau_c,   T[0:c, c] = -tau_c T[0:c, 0:c] (V[:, 0:c]^T v_c),
// with V^T v_c = G[:, c], G = V^T V (for a TSQRT block V = [I; V2] the identity part contributes
// nothing off the diagonal, so G = V2^T V2 there).  One warp, lane = reflector c:
//   vsw   the block in the 128B-swizzled TMA tile layout (reflector j = line j); UNIT: a raw GEQRT
//         tile (R above the diagonal) -- the implied zeros / unit diagonal are substituted;
//   gt    32x32 floats of scratch;
//   ops   six swizzled output tiles, the B operands of the sweeps' three products already split into
//         their 3xTF32 halves (hi = tf32(x), lo = x - hi): V hi, V lo (line c = reflector c), T hi, T lo
//         (column-major: line c = column c of T), V^T hi, V^T lo (line r = tile row r, entry j = V(r, j)).
__device__ __forceinline__ float tf32_round(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u); }
constexpr int OPS_V = 0, OPS_T = 2, OPS_VT = 4;      // tile index of the hi half; lo follows
template <bool UNIT>
__device__ __forceinline__ void build_t_warp(const float *vsw, float tau, float *gt, float *ops, int lane)
{
    float v[TB];
    load_col_sw(v, vsw, lane);
    if (UNIT) {
#pragma unroll
        for (int r = 0; r < TB; ++r) v[r] = (r > lane) ? v[r] : ((r == lane) ? 1.0f : 0.f);
    }
    {   // V (line = lane) and V^T (element lane of every line)
        float *vh = ops + OPS_V * TB * TB, *vl = vh + TB * TB, *th = ops + OPS_VT * TB * TB, *tl = th + TB * TB;
        char *bh = reinterpret_cast<char *>(vh) + lane * 128, *bl = reinterpret_cast<char *>(vl) + lane * 128;
        const int sw = (lane & 7) << 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 h = make_float4(tf32_round(v[4 * i]), tf32_round(v[4 * i + 1]), tf32_round(v[4 * i + 2]), tf32_round(v[4 * i + 3]));
            *reinterpret_cast<float4 *>(bh + ((i << 4) ^ sw)) = h;
            *reinterpret_cast<float4 *>(bl + ((i << 4) ^ sw)) = make_float4(v[4 * i] - h.x, v[4 * i + 1] - h.y, v[4 * i + 2] - h.z, v[4 * i + 3] - h.w);
        }
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const int off = r * TB + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3));
            const float h = tf32_round(v[r]);
            th[off] = h;
            tl[off] = v[r] - h;
        }
    }
    // row `lane` of G^T: G[j][lane] = v_j . v_lane, element j of the row in chunk (j >> 2) ^ (lane & 7)
    {
        float *grow = gt + lane * TB;
#pragma unroll 2
        for (int j = 0; j < TB; ++j) {
            float x[TB];
            const char *base = reinterpret_cast<const char *>(vsw) + j * 128;
            const int sw = (j & 7) << 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 t4 = *reinterpret_cast<const float4 *>(base + ((i << 4) ^ sw));
                x[4 * i] = t4.x; x[4 * i + 1] = t4.y; x[4 * i + 2] = t4.z; x[4 * i + 3] = t4.w;
            }
            if (UNIT) {
#pragma unroll
                for (int r = 0; r < TB; ++r) x[r] = (r > j) ? x[r] : ((r == j) ? 1.0f : 0.f);
            }
            grow[(((j >> 2) ^ (lane & 7)) << 2) | (j & 3)] = dot32(v, x);
        }
    }
    __syncwarp();
    // row `lane` of T, column by column (four partial sums keep the FMA chain short)
    float *toh = ops + OPS_T * TB * TB, *tol = toh + TB * TB;
    float t[TB];
#pragma unroll
    for (int c = 0; c < TB; ++c) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const char *grow = reinterpret_cast<const char *>(gt) + c * 128;      // G[0..31][c]
        const int sw = (c & 7) << 4;
#pragma unroll
        for (int j4 = 0; j4 < c; j4 += 4) {
            const float4 g4 = *reinterpret_cast<const float4 *>(grow + ((j4 << 2) ^ sw));
            a0 = fmaf(t[j4], g4.x, a0);
            if (j4 + 1 < c) a1 = fmaf(t[j4 + 1], g4.y, a1);
            if (j4 + 2 < c) a2 = fmaf(t[j4 + 2], g4.z, a2);
            if (j4 + 3 < c) a3 = fmaf(t[j4 + 3], g4.w, a3);
        }
        const float tc_ = __shfl_sync(FULL, tau, c);
        t[c] = (c == lane) ? tc_ : ((c > lane) ? -tc_ * ((a0 + a1) + (a2 + a3)) : 0.f);
        const int off = c * TB + ((((lane >> 2) ^ (c & 7)) << 2) | (lane & 3));
        const float h = tf32_round(t[c]);
        toh[off] = h;
        tol[off] = t[c] - h;
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
