// qr_persistent_tc.cuh -- tensor-core sweep engine of the persistent kernel (included by
// qr_persistent.cu).  Same scheduler, tickets, panel warps and output layout as the SIMT engine;
// only the way a sweep S(k,g) applies a reflector block to a tile row changes.
//
// For one block of 32 reflectors V (tau, 32x32 tile; V = [I; V2] for TSMQR, unit-lower V1 for
// UNMQR) and 128 matrix columns (one thread each, "thread-per-column" like the SIMT engine):
//     Z = C_bot^T V              tcgen05.mma, M = 128 columns, N = 32, K = 32      (product 1)
//     s = c_top + z              per thread (c_top = the resident tile row k, 32 registers)
//     W = S T                    tcgen05.mma; T = (diag(1/tau) + striu(V^T V))^-1   (product 2)
//     c_top -= w                 per thread
//     C_bot^T -= W V^T           tcgen05.mma onto an accumulator preloaded with C_bot^T (product 3)
// which is the reference's reflector-by-reflector update (qrdecomp.c:736-762) in compact-WY form:
// w_c = tau_c (s_c - sum_{j<c} (v_j' v_c) w_j).  Every product runs as 3xTF32 (hi*hi + hi*lo +
// lo*hi, fp32 accumulation in TMEM), so results stay at float32 accuracy.  V^T V itself is a
// tensor-core product (G); a helper warp inverts the 32x32 triangle in fp32.
// A sweep covers 8 tile columns = 2 blocks of 128 columns; 8 consumer warps, 1 column per thread.
#pragma once
#include "qr_umma.cuh"

namespace qrb200 {

template <>
struct Cfg<3> {
    static constexpr int G = 8;
    static constexpr int NS = 2;
    static constexpr int STAGE_BYTES = G * TILE_BYTES + TILE_BYTES + 1024;      // A tiles (hi in place), V (hi in place), tau
    static constexpr int OFF_LO = NS * STAGE_BYTES;                             // A lo (one buffer)
    static constexpr int OFF_VX = OFF_LO + G * TILE_BYTES;                      // per stage: v_lo, vt_hi, vt_lo, t_hi, t_lo
    static constexpr int VX_BYTES = 5 * TILE_BYTES;
    static constexpr int OFF_SW = OFF_VX + NS * VX_BYTES;                       // S / W operand: hi, lo (256 lines each)
    static constexpr int OFF_GS = OFF_SW + 2 * G * TILE_BYTES;                  // helper scratch: V^T V
    static constexpr int SWEEP_END = OFF_GS + TILE_BYTES;
    static constexpr int PANEL_BYTES = 8 * TILE_BYTES + 5120;
    static constexpr int OFF_PANEL = 0;                                         // role split only: overlays the stages
    static constexpr int OFF_BAR = SWEEP_END;
    static constexpr int SMEM_BYTES = OFF_BAR + 512;
    static constexpr bool BUILD_T = false;
};

namespace tc {
using C = Cfg<3>;
constexpr int W_HELPER = 8, W_LOAD = 9, W_STORE = 10, W_MMA = 11, W_PANEL0 = 12;
constexpr int NTHREADS_TC = (W_PANEL0 + NPW) * 32;
constexpr int SWEEP_THREADS_TC = W_PANEL0 * 32;
constexpr uint32_t TMEM_COLS = 512;
// TMEM column map (128 lanes x 32 columns each)
__device__ __forceinline__ uint32_t tm_z(int s, int blk) { return (uint32_t)((s * 2 + blk) * 32); }        // 0..127
__device__ __forceinline__ uint32_t tm_p(int s, int blk) { return (uint32_t)(128 + (s * 2 + blk) * 32); }  // 128..255
__device__ __forceinline__ uint32_t tm_w(int blk) { return (uint32_t)(256 + blk * 32); }                   // 256..319
__device__ __forceinline__ uint32_t tm_g(int s) { return (uint32_t)(320 + s * 32); }                       // 320..383

struct Bars {
    // indices into the barrier array at OFF_BAR; pfull of the panel warps keeps its slot 3*NS + 2*pw
    static constexpr int FULL = 0, COMPUTED = 2, EMPTY = 4, PFULL = 6, SPLIT = 6 + 2 * NPW, ZREADY = SPLIT + 2,
                         TREADY = ZREADY + 2, SREADY = TREADY + 2, WREADY = SREADY + 1, WSTORED = WREADY + 1,
                         PREADY = WSTORED + 1, COUNT = PREADY + 1;
};
__device__ __forceinline__ uint64_t *bar(int idx) { return reinterpret_cast<uint64_t *>(g_smem + C::OFF_BAR) + idx; }
__device__ __forceinline__ float *a_tile(int s, int t) { return reinterpret_cast<float *>(g_smem + s * C::STAGE_BYTES + t * TILE_BYTES); }
__device__ __forceinline__ float *v_hi(int s) { return a_tile(s, C::G); }
__device__ __forceinline__ float *tau_s(int s) { return a_tile(s, C::G + 1); }
__device__ __forceinline__ float *a_lo(int t) { return reinterpret_cast<float *>(g_smem + C::OFF_LO + t * TILE_BYTES); }
__device__ __forceinline__ float *vx(int s, int which) { return reinterpret_cast<float *>(g_smem + C::OFF_VX + s * C::VX_BYTES + which * TILE_BYTES); }
enum { V_LO = 0, VT_HI = 1, VT_LO = 2, T_HI = 3, T_LO = 4 };
__device__ __forceinline__ float *sw_hi() { return reinterpret_cast<float *>(g_smem + C::OFF_SW); }
__device__ __forceinline__ float *sw_lo() { return reinterpret_cast<float *>(g_smem + C::OFF_SW + C::G * TILE_BYTES); }
__device__ __forceinline__ float *g_scratch() { return reinterpret_cast<float *>(g_smem + C::OFF_GS); }

// write a 32-float line (one matrix column's coefficients) as hi / lo operands, swizzled
__device__ __forceinline__ void store_line_split(float *tile_hi, float *tile_lo, int line, const float (&x)[TB])
{
    char *bh = reinterpret_cast<char *>(tile_hi) + line * 128;
    char *bl = reinterpret_cast<char *>(tile_lo) + line * 128;
    const int sw = (line & 7) << 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 h = make_float4(tf32_hi(x[4 * i]), tf32_hi(x[4 * i + 1]), tf32_hi(x[4 * i + 2]), tf32_hi(x[4 * i + 3]));
        *reinterpret_cast<float4 *>(bh + ((i << 4) ^ sw)) = h;
        *reinterpret_cast<float4 *>(bl + ((i << 4) ^ sw)) =
            make_float4(x[4 * i] - h.x, x[4 * i + 1] - h.y, x[4 * i + 2] - h.z, x[4 * i + 3] - h.w);
    }
}

// three 3xTF32 passes of D[128 x 32] (+)= A[128 lines x 32] * B[32 lines x 32]^T, all K-major / SW128
__device__ __forceinline__ void mma3(uint32_t d, const float *a_h, const float *a_l, const float *b_h, const float *b_l,
                                     uint32_t idesc, bool accumulate)
{
    const float *As[3] = {a_h, a_h, a_l};
    const float *Bs[3] = {b_h, b_l, b_h};
    bool acc = accumulate;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            umma_tf32(d, umma_desc_sw128(As[pass], kk * 32), umma_desc_sw128(Bs[pass], kk * 32), idesc, acc);
            acc = true;
        }
}

// Round / row counters of a role; barrier parities derive from them.  A sweep is nrows reflector rounds plus
// one write-back round.  FULL / COMPUTED / EMPTY flip every time their stage is used (every other round);
// SPLIT / ZREADY / TREADY only in reflector rounds (u[stage] counts those); SREADY .. PREADY once per
// reflector round (r).
struct RowCount {
    int n = 0, r = 0;
    int u[2] = {0, 0};
    __device__ __forceinline__ int stage() const { return n & 1; }
    __device__ __forceinline__ uint32_t ring_par() const { return (n >> 1) & 1; }
    __device__ __forceinline__ void reflector_round_done() { ++u[n & 1]; ++r; ++n; }
    __device__ __forceinline__ void writeback_round_done() { ++n; }
};

// ------------------------------------------------------------------ roles of a sweep
__device__ void loader(const Params &P, const CUtensorMap *mapA, const CUtensorMap *mapVs, int k, int g, RowCount &rc,
                       const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int p = P.p, nrows = p - k;
    const int jt = g * C::G + lane;
    const int kla = k + (lookahead(P, k, C::G) ? 1 : 0);
    const bool active = lane < C::G && jt > kla && jt < P.q;
    const unsigned amask = __ballot_sync(FULL, active);
    const uint32_t bytes = __popc(amask) * TILE_BYTES + TILE_BYTES + 128;
    const int owner = (k / C::G) % P.world;
    const bool remote = owner != P.rank;
    const CUtensorMap *mapV = mapVs + owner;
    const float *tau_src = P.peer_tauc[owner] + (size_t)k * p * TB;
    const int *flag = active ? P.prog + (size_t)jt * p : (lane == C::G ? P.peer_prog[owner] + (size_t)k * p : nullptr);
    const bool flag_remote = remote && lane == C::G;
    const int target = active ? k : k + 1;
    for (int idx = 0; idx <= nrows; ++idx, ++rc.n) {
        const int s = rc.n & 1;
        mbar_wait_wd(bar(Bars::EMPTY + s), ((rc.n >> 1) & 1) ^ 1, wd);
        if (idx < nrows) {
            const int row = k + idx;
            unsigned spins = 0;
            for (;;) {
                bool ok = flag ? (flag_remote ? ld_relaxed_sys(flag + row) : ld_relaxed(flag + row)) >= target : true;
                if (__all_sync(FULL, ok)) break;
                __nanosleep(64);
                wd.check(spins);
            }
            fence_acq_rel(remote);
            fence_proxy_async();
            if (lane == 0) mbar_arrive_expect_tx(bar(Bars::FULL + s), bytes);
            __syncwarp();
            if (active) tma_load_2d(a_tile(s, lane), mapA, row * TB, jt * TB, bar(Bars::FULL + s));
            if (lane == C::G) tma_load_2d(v_hi(s), mapV, row * TB, k * TB, bar(Bars::FULL + s));
            if (lane == C::G + 1) bulk_g2s(tau_s(s), tau_src + (size_t)row * TB, 128, bar(Bars::FULL + s));
        } else if (lane == 0) {
            mbar_arrive(bar(Bars::FULL + s));
        }
    }
}

__device__ void storer(const Params &P, const CUtensorMap *mapA, int k, int g, RowCount &rc, const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int p = P.p, nrows = p - k;
    const int jt = g * C::G + lane;
    const int kla = k + (lookahead(P, k, C::G) ? 1 : 0);
    const bool active = lane < C::G && jt > kla && jt < P.q;
    for (int idx = 0; idx <= nrows; ++idx, ++rc.n) {
        const int s = rc.n & 1;
        mbar_wait_wd(bar(Bars::COMPUTED + s), (rc.n >> 1) & 1, wd);
        if (idx > 0) {
            const int row = idx < nrows ? k + idx : k;
            if (active) {
                tma_store_2d(mapA, row * TB, jt * TB, a_tile(s, lane));
                bulk_commit();
                bulk_wait_all();
                __threadfence();
                st_release(P.prog + (size_t)jt * p + row, k + 1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Bars::EMPTY + s));
    }
}

// helper warp: split V (and its transpose) into operands, then T = (diag(1/tau) + striu(V^T V))^-1
__device__ void helper(const Params &P, int k, uint32_t tmem, RowCount &rc, const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int nrows = P.p - k;
    float *gt = g_scratch();                                       // (V^T V) transposed: gt[c*32 + j]
    for (int idx = 0; idx < nrows; ++idx) {
        const int s = rc.stage();
        mbar_wait_wd(bar(Bars::FULL + s), rc.ring_par(), wd);
        {   // reflector column `lane`: hi in place, lo, and the transposed copies
            float v[TB];
            load_col_sw(v, v_hi(s), lane);
            if (idx == 0) {
#pragma unroll
                for (int r = 0; r < TB; ++r) v[r] = (r > lane) ? v[r] : ((r == lane) ? 1.0f : 0.f);   // unit lower V1
            }
            float *vth = vx(s, VT_HI), *vtl = vx(s, VT_LO);
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const float h = tf32_hi(v[r]);
                const int off = r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3));
                vth[off] = h;
                vtl[off] = v[r] - h;
            }
            store_line_split(v_hi(s), vx(s, V_LO), lane, v);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar(Bars::SPLIT + s));
        // V^T V comes back from the tensor core together with Z of this row
        mbar_wait_wd(bar(Bars::ZREADY + s), rc.u[s] & 1, wd);
        tc_fence_after();
        float grow[TB];
        tmem_ld_32x32(grow, tmem + tm_g(s));                        // lane j: (V^T V)[j][0..31]
#pragma unroll
        for (int c = 0; c < TB; ++c) gt[c * TB + lane] = grow[c];
        const float *ts = tau_s(s);
        __syncwarp();
        // row `lane` of T:  T[r][r] = tau_r,  T[r][c] = -tau_c * sum_{j=r}^{c-1} T[r][j] G[j][c]   (c > r)
        float t[TB];
        float *th = vx(s, T_HI), *tl = vx(s, T_LO);
#pragma unroll
        for (int c = 0; c < TB; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < c; j4 += 4) {                     // G[j][c], j < c: contiguous in gt
                const float4 g4 = *reinterpret_cast<const float4 *>(gt + c * TB + j4);
                acc = fmaf(t[j4], g4.x, acc);
                if (j4 + 1 < c) acc = fmaf(t[j4 + 1], g4.y, acc);
                if (j4 + 2 < c) acc = fmaf(t[j4 + 2], g4.z, acc);
                if (j4 + 3 < c) acc = fmaf(t[j4 + 3], g4.w, acc);
            }
            const float tc_ = ts[c];
            t[c] = (c == lane) ? tc_ : ((c > lane) ? -tc_ * acc : 0.f);
            // B operand of product 2: line c = column c of T; this lane writes element `lane` of it
            const float h = tf32_hi(t[c]);
            const int off = c * 32 + ((((lane >> 2) ^ (c & 7)) << 2) | (lane & 3));
            th[off] = h;
            tl[off] = t[c] - h;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Bars::TREADY + s));
        rc.reflector_round_done();
    }
    rc.writeback_round_done();
}

// the MMA issuer: lane 0 of its warp; everything else in the CTA talks to it through mbarriers
__device__ void mma_issuer(const Params &P, int k, uint32_t tmem, RowCount &rc, const Watch &wd)
{
    const int nrows = P.p - k;
    const bool me = (threadIdx.x & 31) == 0;
    const uint32_t idesc = umma_idesc_tf32(128, 32, false, false);
    const uint32_t idesc_neg = umma_idesc_tf32(128, 32, false, false, true);
    auto issue_gz = [&](int s, int upar) {       // V^T V and Z of a row staged in s
        mbar_wait_wd(bar(Bars::SPLIT + s), upar, wd);
        tc_fence_after();
        if (me) {
            mma3(tmem + tm_g(s), v_hi(s), vx(s, V_LO), v_hi(s), vx(s, V_LO), idesc, false);   // rows >= 32 of D: unused
            for (int blk = 0; blk < 2; ++blk)
                mma3(tmem + tm_z(s, blk), a_tile(s, blk * 4), a_lo(blk * 4), v_hi(s), vx(s, V_LO), idesc, false);
            umma_commit(bar(Bars::ZREADY + s));
        }
        __syncwarp();
    };
    issue_gz(rc.stage(), rc.u[rc.stage()] & 1);
    for (int idx = 0; idx < nrows; ++idx) {
        const int s = rc.stage();
        // product 2: W = S T
        mbar_wait_wd(bar(Bars::SREADY), rc.r & 1, wd);
        mbar_wait_wd(bar(Bars::TREADY + s), rc.u[s] & 1, wd);
        tc_fence_after();
        if (me) {
            for (int blk = 0; blk < 2; ++blk)
                mma3(tmem + tm_w(blk), sw_hi() + blk * 4 * TB * TB, sw_lo() + blk * 4 * TB * TB, vx(s, T_HI), vx(s, T_LO),
                     idesc, false);
            umma_commit(bar(Bars::WREADY));
        }
        __syncwarp();
        // next row's first products overlap the consumers' work on this one
        if (idx + 1 < nrows) issue_gz(s ^ 1, rc.u[s ^ 1] & 1);
        // product 3: C_bot^T -= W V^T  (accumulator preloaded with C_bot^T by the consumers)
        mbar_wait_wd(bar(Bars::WSTORED), rc.r & 1, wd);
        tc_fence_after();
        if (me) {
            for (int blk = 0; blk < 2; ++blk)
                mma3(tmem + tm_p(s, blk), sw_hi() + blk * 4 * TB * TB, sw_lo() + blk * 4 * TB * TB, vx(s, VT_HI),
                     vx(s, VT_LO), idesc_neg, true);
            umma_commit(bar(Bars::PREADY));
        }
        __syncwarp();
        rc.reflector_round_done();
    }
    rc.writeback_round_done();
}

// consumer thread: one matrix column.  warp w: tile column w of the sweep, block w/4, TMEM lanes 32*(w%4)..
__device__ void consumer(const Params &P, int k, int g, uint32_t tmem, RowCount &rc, const Watch &wd)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = w >> 2;
    const int nrows = P.p - k;
    const int jt = g * C::G + w;
    const bool act = jt > k + (lookahead(P, k, C::G) ? 1 : 0) && jt < P.q;
    const uint32_t lane_base = (uint32_t)((w & 3) * 32) << 16;
    const int line = (w & 3) * 32 + lane;                      // line of this column inside its 128-line block
    float ctop[TB];
#pragma unroll
    for (int r = 0; r < TB; ++r) ctop[r] = 0.f;

    auto split_row = [&](int s, uint32_t ring_par) {   // operands of product 1 + accumulator preload of product 3
        mbar_wait_wd(bar(Bars::FULL + s), ring_par, wd);
        if (act) {
            float x[TB];
            load_col_sw(x, a_tile(s, w), lane);
            tmem_st_32x32(tmem + lane_base + tm_p(s, blk), x);
            store_line_split(a_tile(s, w), a_lo(w), lane, x);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar(Bars::SPLIT + s));
    };

    split_row(rc.stage(), rc.ring_par());
    for (int idx = 0; idx <= nrows; ++idx) {
        const int s = rc.stage();
        const uint32_t rowpar = rc.r & 1;
        if (idx < nrows) {
            // product 1 result: z = V^T c_bot for this column
            mbar_wait_wd(bar(Bars::ZREADY + s), rc.u[s] & 1, wd);
            tc_fence_after();
            float sv[TB];
            if (act) {
                tmem_ld_32x32(sv, tmem + lane_base + tm_z(s, blk));
                if (idx > 0) {
#pragma unroll
                    for (int r = 0; r < TB; ++r) sv[r] += ctop[r];
                }
                store_line_split(sw_hi() + blk * 4 * TB * TB, sw_lo() + blk * 4 * TB * TB, line, sv);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar(Bars::SREADY));
            // overlap: operands of the next row (A lo is free: product 1 of this row has completed)
            if (idx + 1 < nrows) split_row(s ^ 1, ((rc.n + 1) >> 1) & 1);
            // product 2 result: w
            mbar_wait_wd(bar(Bars::WREADY), rowpar, wd);
            tc_fence_after();
            if (act) {
                tmem_ld_32x32(sv, tmem + lane_base + tm_w(blk));
                if (idx > 0) {
#pragma unroll
                    for (int r = 0; r < TB; ++r) ctop[r] -= sv[r];
                }
                store_line_split(sw_hi() + blk * 4 * TB * TB, sw_lo() + blk * 4 * TB * TB, line, sv);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar(Bars::WSTORED));
            // product 3 result: the updated tile column
            mbar_wait_wd(bar(Bars::PREADY), rowpar, wd);
            tc_fence_after();
            if (act) {
                tmem_ld_32x32(sv, tmem + lane_base + tm_p(s, blk));
                if (idx == 0) {
#pragma unroll
                    for (int r = 0; r < TB; ++r) ctop[r] = sv[r];       // UNMQR: the result IS the resident tile row k
                } else {
                    store_col_sw(a_tile(s, w), lane, sv);
                }
            }
        } else {
            // write-back round: the resident tile row k goes out through stage s
            mbar_wait_wd(bar(Bars::FULL + s), rc.ring_par(), wd);
            if (act) store_col_sw(a_tile(s, w), lane, ctop);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar(Bars::COMPUTED + s));
        if (idx < nrows) rc.reflector_round_done(); else rc.writeback_round_done();
    }
}

}  // namespace tc

__global__ void __launch_bounds__(tc::NTHREADS_TC, 1) tiledqr_persistent_tc(const __grid_constant__ KernelArgs KA)
{
    using namespace tc;
    const int vr = KA.nvirt > 1 ? (int)(blockIdx.x % KA.nvirt) : 0;
    const int bid = KA.nvirt > 1 ? (int)(blockIdx.x / KA.nvirt) : (int)blockIdx.x;
    const Params &P = KA.rc[vr].P;
    const CUtensorMap &mapA = KA.rc[vr].mapA;
    const CUtensorMap *mapVs = KA.rc[vr].mapV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Watch wd{gtimer(), P.watchdog_ns};
    const bool panel_cta = bid < P.npanel_ctas;
    int *s_task = reinterpret_cast<int *>(bar(Bars::COUNT));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_task + 4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(Bars::FULL + s), 1);
            mbar_init(bar(Bars::COMPUTED + s), 256);
            mbar_init(bar(Bars::EMPTY + s), 1);
            mbar_init(bar(Bars::SPLIT + s), 256 + 32);
            mbar_init(bar(Bars::ZREADY + s), 1);
            mbar_init(bar(Bars::TREADY + s), 1);
        }
        mbar_init(bar(Bars::SREADY), 256);
        mbar_init(bar(Bars::WREADY), 1);
        mbar_init(bar(Bars::WSTORED), 256);
        mbar_init(bar(Bars::PREADY), 1);
        for (int i = 0; i < 2 * NPW; ++i) mbar_init(bar(Bars::PFULL + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mapA);
        for (int r = 0; r < P.world; ++r) tma_prefetch_desc(&mapVs[r]);
    }
    if (warp == W_MMA && !panel_cta) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= W_PANEL0) {
        if (!panel_cta) return;
        uint32_t phbits = 0;
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&P.ctr[2], 1);
            t = __shfl_sync(FULL, t, 0);
            if (t >= P.npanels) break;
            run_panel<3>(P, &mapA, P.panel_tab[t], warp - W_PANEL0, phbits, wd);
        }
        return;
    }
    if (panel_cta) return;

    const uint32_t tmem = *tmem_slot;
    RowCount rc;
    for (;;) {
        named_bar_sync(1, SWEEP_THREADS_TC);
        if (warp == W_LOAD && lane == 0) {
            int t = atomicAdd(&P.ctr[0], 1);
            int2 kg = t < P.nsweeps ? P.sweep_tab[t] : make_int2(-1, 0);
            s_task[0] = kg.x; s_task[1] = kg.y;
        }
        named_bar_sync(1, SWEEP_THREADS_TC);
        const int k = s_task[0], g = s_task[1];
        if (k < 0) break;
        {
            const int j_lo = max(g * C::G, k + 1 + (lookahead(P, k, C::G) ? 1 : 0));
            const int j_hi = min((g + 1) * C::G, P.q);
            if (j_lo >= j_hi) continue;
        }
        if (warp == W_LOAD) loader(P, &mapA, mapVs, k, g, rc, wd);
        else if (warp == W_STORE) storer(P, &mapA, k, g, rc, wd);
        else if (warp == W_HELPER) helper(P, k, tmem, rc, wd);
        else if (warp == W_MMA) mma_issuer(P, k, tmem, rc, wd);
        else consumer(P, k, g, tmem, rc, wd);
    }
    tc_fence_before();
    named_bar_sync(1, SWEEP_THREADS_TC);
    if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace qrb200
